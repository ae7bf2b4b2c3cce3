"""The DEVICE_GPU TensorFlow op wrappers (integration/tf_ops_gpu.cc) cannot be built here (no
TensorFlow in the image): they are type-checked against include/ssnt_tts_c.h and a stand-in for the few TensorFlow
types they touch (integration/tf_mock), and checked to register a GPU kernel for each of the reference's seven ops."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "integration", "tf_ops_gpu.cc")


def test_tf_gpu_ops_type_check_against_the_c_header():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "integration", "tf_mock"),
                        "-I", os.path.join(ROOT, "include"), SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_reference_op_gets_a_gpu_kernel():
    text = open(SRC).read()
    registered = set(re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)\.Device\(tf::DEVICE_GPU\)', text))
    # REGISTER_OP names in ssnt-tts-tensorflow/src/*.cc
    reference_ops = {"SSNTBeamSearchDecode", "SSNTExtractBestBeamBranch", "SSNTV2BeamSearchDecode", "SSNTOrderBeamBranch",
                     "SSNTUpsampleSourceIndexes", "ToneLatentBeamSearchDecode", "ToneLatentLevenshteinEditDistance"}
    assert reference_ops <= registered
    assert {"SSNTForwardBackward", "SSNTForwardBackwardLogits"} <= registered
