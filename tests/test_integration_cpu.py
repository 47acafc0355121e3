"""The plugin of INTEGRATION.md (integration/b200_ops.cc: the hot-path operators and their neighbours written against the REFERENCE's own headers --
operator.h, crop_attr.h, resize_attr.h, resampling_attr.h -- plus AudioResample / NonsilentRegion derived from the reference's own
operator bases (audio::ResampleBase, NonsilenceOperator: its argument handling and shape inference, our RunImpl), all over the C-ABI
of include/dali_b200.h) must at least compile against those headers: `g++ -std=c++20 -fsyntax-only` where /root/reference exists (it
does not on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dali")), reason="needs the reference tree")
def test_plugin_compiles_against_reference_headers():
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-w", "-I" + REF, "-I" + os.path.join(REF, "include"), "-I/usr/local/cuda/include",
           "-I" + os.path.join(REF, "third_party/boost/preprocessor/include"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "b200_ops.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
    src = open(os.path.join(ROOT, "integration", "b200_ops.cc")).read()
    for op in ("decoders__Image", "Resize", "CropMirrorNormalize", "WarpAffine", "Hsv", "ColorSpaceConversion", "Spectrogram", "MelFilterBank",
               "AudioResample", "NonsilentRegion", "decoders__ImageCrop", "decoders__ImageRandomCrop", "ColorTwist", "BrightnessContrast",
               "Flip", "Crop", "RandomResizedCrop", "ToDecibels", "MFCC", "decoders__ImageSlice", "Slice", "Rotate"):
        assert f"DALI_REGISTER_OPERATOR(b200__{op}," in src, op


def test_c_header_is_plain_c():
    """include/dali_b200.h is the drop-in boundary: it must compile as C (no C++ types in any signature)."""
    code = '#include "dali_b200.h"\nint main(void) { return dalib200GetVersion == 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "include"), "-x", "c", "-"], input=code,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
