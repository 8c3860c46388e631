"""CPU: the C-ABI library loads and exports exactly what include/sidgpu.h declares; without a GPU
every entry point that needs one fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sidgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sidgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sid_b200 import _lib, build
    build.build_libsidgpu()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "sid_b200", "libsidgpu.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "orc_" not in out and "ref_" not in out
    for path, _, files in os.walk(os.path.join(ROOT, "sid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inl")):
                text = open(os.path.join(path, f)).read()
                assert "liboracle" not in text and "oracle/" not in text and "oracle_py" not in text, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sid_b200
    with pytest.raises(sid_b200.SidGpuError) as e:
        sid_b200.Context()
    assert "no CPU fallback" in str(e.value)


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirror of the ABI's structs has the C compiler's sizes and offsets."""
    import ctypes
    import subprocess
    from sid_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "sidgpu.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu\\n", sizeof(sidgpu_params), offsetof(sidgpu_params, het_only), offsetof(sidgpu_params, fit_nd), sizeof(sidgpu_config));\n'
        '  printf("%zu %zu %zu\\n", sizeof(sidgpu_sites_view), sizeof(sidgpu_unique_view), sizeof(sidgpu_fit));\n'
        '  printf("%zu %zu %zu\\n", sizeof(sidgpu_bgzf_block), offsetof(sidgpu_bgzf_block, isize), offsetof(sidgpu_bgzf_block, crc));\n'
        '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, check=True, text=True).stdout.split()
    got = [int(x) for x in out]
    want = [ctypes.sizeof(_lib.Params), _lib.Params.het_only.offset, _lib.Params.fit_nd.offset, ctypes.sizeof(_lib.Config),
            ctypes.sizeof(_lib.SitesView), ctypes.sizeof(_lib.UniqueView), ctypes.sizeof(_lib.Fit),
            ctypes.sizeof(_lib.BgzfBlock), _lib.BgzfBlock.isize.offset, _lib.BgzfBlock.crc.offset]
    assert got == want
