"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/promptir_b200.h declares, and the
ctypes mirrors in promptir_b200/_lib.py have the same size/offsets as the C structs (checked with gcc)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from promptir_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "promptir_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(pir_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pir_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported


def test_host_only_entry_points(lib):
    assert lib.pir_abi_version() == 1
    assert 1 <= lib.pir_mdta_splits(16, 65536, 96) <= 64 and lib.pir_mdta_splits(1, 64, 384) == 1
    assert lib.pir_mdta_ws_floats(2, 48, 3) == 2 * 3 * (48 * 48 + 96) + 2 * 48 * 48
    assert lib.pir_mdta_finalize_kernels(384, 8) in (1, 2) and lib.pir_mdta_finalize_kernels(704, 1) == 2
    assert lib.pir_prompt_ws_floats(2, 1024, 384) == 2 * 4 * 384
    # argument validation happens before any CUDA call: a null descriptor is rejected with a message
    assert lib.pir_gemm(None, None) == -1 and b"null" in lib.pir_last_error()


def test_ctypes_structs_match_the_header(tmp_path):
    names = ["PirGemm", "PirDwConv", "PirPwDw", "PirMdta", "PirPrompt", "PirPatchEmbed", "PirLn", "PirWgrad", "PirWgradFin", "PirDwWgrad",
             "PirGateBwd", "PirMdtaBwd", "PirShuffle", "PirPromptBwd", "PirBcastAdd", "PirToNhwc16", "PirOcab", "PirOcabBwd", "PirPackJob"]
    fields = {n: [f[0].rstrip("_") if f[0] == "in_" else f[0] for f in getattr(_lib, n)._fields_] for n in names}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for n in names:
        lines.append(f'printf("{n} %zu", sizeof({n}));')
        for f in fields[n]:
            lines.append(f'printf(" %zu", offsetof({n}, {f}));')
        lines.append('printf("\\n");')
    lines.append("return 0;}")
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        cls = getattr(_lib, parts[0])
        assert ctypes.sizeof(cls) == int(parts[1]), parts[0]
        for (fname, _), off in zip(cls._fields_, parts[2:]):
            assert getattr(cls, fname).offset == int(off), (parts[0], fname)
