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


# cfg/octree-liftt-ctc-lossless-geom-lossy-attrs.yaml (transformType 2) and
# cfg/octree-predt-ctc-lossless-geom-nearlossless-attrs.yaml (transformType 1), colour
def lod_flags(qp=34, transform_type=2, decimator=0, lods=10):
    f = [x for x in enc_flags(qp, transform_type) if not x.startswith("--rahtPredictionSearchRange")
         and not x.startswith("--qpChromaOffset")]
    i = f.index(f"--qp={qp}")
    extra = ["--numberOfNearestNeighborsInPrediction=3", f"--levelOfDetailCount={lods}",
             f"--lodDecimator={decimator}", "--adaptivePredictionThreshold=64", "--qpChromaOffset=0"]
    if decimator:
        extra += ["--lodSamplingPeriod=4", "--lod_neigh_bias=1,1,1"]
    if transform_type == 1:
        extra += ["--intraLodPredictionSkipLayers=0", "--interComponentPredictionEnabled=0",
                  "--predWeightBlending=1"]
    return f[:i] + extra + f[i:]


def encode(binary, ply, out_bin, out_rec, qp=34, flags=None):
    cmd = [binary, f"--uncompressedDataPath={ply}", f"--compressedStreamPath={out_bin}",
           f"--reconstructedDataPath={out_rec}"] + (flags if flags is not None else enc_flags(qp))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout


def decode(binary, in_bin, out_rec):
    cmd = [binary, "--mode=1", f"--compressedStreamPath={in_bin}",
           f"--reconstructedDataPath={out_rec}", "--convertPlyColourspace=1"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout
