"""CPU: the C-ABI library loads and exports every symbol include/rpo_b200.h declares; argument
validation that needs no GPU behaves; the product path refuses CPU tensors."""
import ctypes as C
import os
import re

import pytest
import torch

from rpo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "rpo_b200.h")).read()
    declared = set(re.findall(r"\b(rpo_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_release_library_reads_three_environment_variables(lib):
    """Tuning and fault-injection switches exist only in the diagnostics build (common.cuh::diag_env): the release
    library must contain no other RPO_* variable names than the three documented ones."""
    if os.environ.get("RPO_DIAG") == "1":
        pytest.skip("diagnostics build")
    blob = open(_lib.LIB_PATH, "rb").read()
    names = set(m.decode() for m in re.findall(rb"RPO_[A-Z][A-Z0-9_]{3,}(?=\x00)", blob))
    names -= {n for n in names if n.startswith(("RPO_ERR_", "RPO_OK", "RPO_F", "RPO_BF", "RPO_U8", "RPO_ACT_", "RPO_GEMM_AUTO",
                                                 "RPO_GEMM_SIMT", "RPO_GEMM_TCGEN05", "RPO_PEER_", "RPO_REQUIRE", "RPO_CHECK_",
                                                 "RPO_TRY", "RPO_LAUNCH_"))}
    assert names == {"RPO_NO_PDL", "RPO_SINGLE_STREAM", "RPO_GEMM_DYNAMIC"}, names


def test_version_and_errors(lib):
    assert lib.rpo_version() >= 100
    h = C.c_void_p()
    cfg = _lib.RpoConfig(dtype=7)
    assert lib.rpo_create(C.byref(cfg), C.byref(h)) == -1
    assert b"dtype" in lib.rpo_last_error()
    cfg = _lib.RpoConfig(dtype=_lib.RPO_F16, K=0, n_cls=2, ctx_len=77, embed_dim=512, v_width=768, v_layers=12,
                         v_heads=12, v_patch=16, v_res=224, t_width=512, t_layers=12, t_heads=8, max_batch=2)
    assert lib.rpo_create(C.byref(cfg), C.byref(h)) == -1
    assert b"K should be bigger than 0" in lib.rpo_last_error()  # trainers/rpo.py:47
    with pytest.raises(_lib.RpoError):
        _lib.check(lib.rpo_backward(None, None, None))


def test_no_cpu_fallback():
    from types import SimpleNamespace
    from rpo_b200 import synth
    from rpo_b200.clip_weights import SyntheticCLIP
    from rpo_b200.model import CustomCLIP
    from tests.common import class_tokens
    arch = synth.ARCHS["tiny"]
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=2, PREC="fp16")),
                          INPUT=SimpleNamespace(SIZE=(64, 64)))
    model = CustomCLIP(cfg, ["a", "b"], "a photo of a _.", SyntheticCLIP(synth.make_state_dict(arch, 0), "fp16"),
                       tokens=class_tokens([1, 2]))
    # the surface the reference trainer relies on
    assert [n for n, _ in model.named_parameters()] == ["prompt_learner.text_prompt", "prompt_learner.img_prompt"]
    assert set(model.prompt_learner.state_dict()) == {"text_prompt", "img_prompt"}
    assert model.prompt_learner.text_prompt.shape == (2, 128) and model.prompt_learner.text_prompt.dtype == torch.float16
    with pytest.raises(_lib.RpoError):
        model(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.int64))


def test_header_is_plain_c_and_a_c_client_links(lib, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a torch-free C program links against the shared
    library and gets the documented status codes (integration/c_client.c)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "rpo_b200.h")], check=True)
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "c_client")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", inc, os.path.join(ROOT, "integration", "c_client.c"),
                    "-L", libdir, "-lrpo_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "argument validation ok" in r.stdout
