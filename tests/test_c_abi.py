"""The C ABI seen from a plain C translation unit (tests/c_abi/abi_smoke.c): compiled with gcc against include/femb200.h
and linked to libfemb200.so -- no ctypes struct mirror involved, so header drift shows up here.
CPU part: it compiles with -Wall -Werror as C99, links, loads the library and agrees on the ABI version.
GPU part: it assembles the reference's ExampleGeneralAssembler tetrahedron and prints the golden 12x12 (example_test.go:72-85)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c")
EXE = os.path.join(ROOT, "tests", "c_abi", "abi_smoke")


def build_abi_smoke(pkg):
    libdir = os.path.dirname(pkg.LIB_PATH)
    newest = max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "femb200.h")), os.path.getmtime(pkg.LIB_PATH))
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + libdir, "-lfemb200",
               "-Wl,-rpath,$ORIGIN/../../go-fem_b200", "-o", EXE]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def test_c_translation_unit_compiles_links_and_loads(pkg):
    exe = build_abi_smoke(pkg)
    p = subprocess.run([exe, "--abi"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    got, want = p.stdout.split()
    assert got == want == "1"


@pytest.mark.gpu
def test_c_program_assembles_the_golden_tetrahedron(pkg):
    exe = build_abi_smoke(pkg)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, f"abi_smoke exit {p.returncode}: {p.stderr}"
    rows = [r.split() for r in p.stdout.strip().splitlines()]
    assert len(rows) == 12 and all(len(r) == 12 for r in rows)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "example_test_outputs.json")))["ExampleGeneralAssembler"]
    assert gold["format"] == "%.5g"
    G = np.array(gold["K"])
    F = np.array([[float(v) for v in r] for r in rows])
    assert np.array_equal(F, G), "the C program's K differs from the reference's Output block (example_test.go:72-85)"
