"""Whole-codec integration harness (TEST INFRASTRUCTURE): runs the reference's
own `tmc3` executable, built by `make -C oracle codec` either unmodified
(oracle/_ref/tmc3_ref) or with tmc3/RAHT.cpp replaced by the product's drop-in
translation unit (oracle/_ref/tmc3_b200), on a synthetic PLY."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tmc3_ref")
B200_BIN = os.path.join(ROOT, "oracle", "_ref", "tmc3_b200")

# cfg/octree-raht-ctc-lossless-geom-lossy-attrs.yaml, rate point r04, colour
def enc_flags(qp=34, transform_type=0):
    return [
    "--mode=0", "--trisoupNodeSizeLog2=0", "--mergeDuplicatedPoints=0",
    "--neighbourAvailBoundaryLog2=8", "--intra_pred_max_node_size_log2=6",
    "--positionQuantizationScale=1", "--inferredDirectCodingMode=1",
    "--rahtPredictionSearchRange=50000", "--maxNumQtBtBeforeOt=4", "--minQtbtSizeLog2=0",
    "--planarEnabled=1", "--planarModeIdcmUse=0", "--convertPlyColourspace=1",
    f"--transformType={transform_type}", f"--qp={qp}", "--qpChromaOffset=-1", "--bitdepth=8",
    "--attrOffset=0", "--attrScale=1", "--attribute=color",
    ]


def write_ply(path, xyz, rgb):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
                "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
                "end_header\n" % len(xyz))
        for p, c in zip(xyz, rgb):
            f.write("%d %d %d %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))


def encode(binary, ply, out_bin, out_rec, qp=34):
    cmd = [binary, f"--uncompressedDataPath={ply}", f"--compressedStreamPath={out_bin}",
           f"--reconstructedDataPath={out_rec}"] + enc_flags(qp)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout


def decode(binary, in_bin, out_rec):
    cmd = [binary, "--mode=1", f"--compressedStreamPath={in_bin}",
           f"--reconstructedDataPath={out_rec}", "--convertPlyColourspace=1"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout
