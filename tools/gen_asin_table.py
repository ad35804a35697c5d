"""Generates the arcsine table of the fixed-point atan2 (G-PCC `iatan2`):
entry i = round(asin(i / 512) * 2^20), i = 0..362, and entry 363 repeats 362
(interpolation guard).  The product embeds the output in
mpeg-pcc-tmc13_b200/csrc/spherical.cuh; tests/test_oracle_vs_reference.py pins
it against the compiled reference."""
import math

vals = [int(math.floor(math.asin(i / 512.0) * (1 << 20) + 0.5)) for i in range(363)]
vals.append(vals[-1])
rows = []
for i in range(0, len(vals), 10):
    rows.append("  " + ", ".join(str(v) for v in vals[i:i + 10]))
print("#define PCC_ASIN_Q20_VALUES \\\n" + ", \\\n".join(rows))
