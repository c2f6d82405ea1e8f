"""The C++ drop-in (hyperpose::parser::paf / hyperpose::dnn::tensorrt over the C ABI):
  * CPU: compiles and links against the reference's UNCHANGED public headers (when /root/reference exists);
  * GPU: the prebuilt example binary runs the reference's operator-API sequence end to end."""
import os
import subprocess

import pytest

from hyperpose_b200 import build as hb, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/include/hyperpose"), reason="reference headers absent")
def test_dropin_compiles_against_unchanged_reference_headers():
    exe = os.path.join(ROOT, "examples", "operator_api_b200")
    if os.path.exists(exe):
        os.remove(exe)
    assert hb.build_cpp_example() == exe and os.path.exists(exe)
    syms = subprocess.run(["nm", "-C", "--defined-only", exe], capture_output=True, text=True).stdout
    for want in ["hyperpose::parser::paf::process(", "hyperpose::parser::paf::paf(float, float, cv::Size)",
                 "hyperpose::parser::paf::set_conf_thresh(float)", "hyperpose::dnn::tensorrt::inference(std::vector<cv::Mat",
                 "hyperpose::dnn::tensorrt::inference(std::vector<float", "hyperpose::dnn::tensorrt::save(", "hyperpose::parser::pifpaf::process(",
                 "hyperpose::parser::pose_proposal::process(", "hyperpose::parser::pose_proposal::set_nms_thresh(float)"]:
        assert want in syms, want


@pytest.mark.gpu
def test_cpp_example_runs_operator_api_sequence(tmp_path):
    exe = hb.build_cpp_example()
    if exe is None:
        pytest.skip("example binary not built (needs the reference headers at build time)")
    pack = tmp_path / "tiny.pack"
    pack.write_bytes(models.tiny_test_net(0).to_pack())
    saved = tmp_path / "saved.pack"
    r = subprocess.run([exe, str(pack), "96", "64", "3", str(saved)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert saved.read_bytes() == pack.read_bytes()      # tensorrt::save re-emits the pack
    assert "conf:[19, 32, 48, ]" in r.stdout and "paf:[38, 32, 48, ]" in r.stdout
    assert "3 images got processed" in r.stdout
