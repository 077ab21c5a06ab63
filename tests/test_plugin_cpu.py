"""CPU: run-time plugin registration (genpf_model_compile, SURVEY 8b / north star "models registered as device
plugins").  NVRTC cross-compiles for sm_100a without a GPU, so compile / registry / image round trip are checked
here; the compiled kernels run in tests/test_gpu_plugin.py."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "tests", "plugins", "my_lingauss.cu")).read()


@pytest.fixture(scope="module")
def plug():
    import genpf_b200 as g
    g.load()
    try:
        return g, g.DeviceModel.from_source("cpu_lingauss", SRC, "MyLinGauss", fields={"x": 0})
    except g.GenPFErrorException as e:
        if "libnvrtc" in str(e):
            pytest.skip("NVRTC is not installed")
        raise


def test_compile_and_registry(plug):
    g, m = plug
    assert m.model_id >= 100 and (m.n_f64, m.n_u8, m.n_params, m.n_aux) == (1, 0, 8, 0)
    assert m.has_translator and not m.has_proposal
    again = g.DeviceModel("cpu_lingauss")  # found by name like a built-in
    assert again.model_id == m.model_id
    om = g.DeviceModel("object_motion")
    assert om.has_proposal and not om.has_translator and om.model_id < 100


def test_compile_error_carries_the_log(plug):
    g, _ = plug
    with pytest.raises(g.GenPFErrorException, match="does not compile"):
        g.DeviceModel.from_source("broken", "struct X { int y }", "X")
    with pytest.raises(g.GenPFErrorException):  # compiles, but is not a plugin: no NF / transition ...
        g.DeviceModel.from_source("notaplugin", "struct Y { int y; };", "Y")


def test_image_round_trip(plug):
    g, m = plug
    img = m.export_image()
    assert img[:8] == b"GENPFPLG" and len(img) > 10_000
    m2 = g.DeviceModel.from_image(img.replace(b"cpu_lingauss", b"cpu_lingausz"), fields={"x": 0})
    assert m2.name == "cpu_lingausz" and m2.model_id != m.model_id
    assert (m2.n_f64, m2.n_u8, m2.n_params, m2.has_translator) == (1, 0, 8, True)
    with pytest.raises(g.GenPFError):
        g.DeviceModel.from_image(b"not an image" * 10)


def test_embedded_headers_match_the_tree(plug):
    """The kernel headers handed to NVRTC are the ones the library itself was built from."""
    import ctypes as C
    g, _ = plug
    lib = g.load()
    seen = {}
    for i in range(16):
        name, text = C.c_char_p(), C.c_char_p()
        if lib.genpf_model_plugin_sources(i, C.byref(name), C.byref(text)) != 0:
            break
        seen[name.value.decode()] = text.value.decode()
    assert {"common.cuh", "models.cuh", "kernels.cuh", "filter.cuh", "fused.cuh"} <= set(seen)
    csrc = os.path.join(ROOT, "genparticlefilters.jl_b200", "csrc")
    for k, v in seen.items():
        assert v == open(os.path.join(csrc, k)).read(), k
