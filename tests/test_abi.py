"""CPU: the C-ABI library loads and exports every symbol include/psim_b200.h declares; without a CUDA
device context creation fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

from helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psim_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from particlesim_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libpsim_b200.so does not export {s}"
    # and the Python binding covers the whole header
    assert set(syms) == set(_lib.SIGNATURES)


def test_struct_layouts_match_the_header():
    from particlesim_b200 import _lib
    assert C.sizeof(_lib.Config) == 64
    assert _lib.SPECIES_DTYPE.itemsize == 48 and _lib.NODE_DTYPE.itemsize == 64
    cfg = _lib.default_config()
    assert (cfg.theta, cfg.epsilon, cfg.leaf_capacity, cfg.thread_capacity) == (1.0, 2.0, 1, 1024)
    assert cfg.lj_force_max == 200.0 and cfg.collision_passes == 7 and cfg.parity_mode == 1


def test_species_table_matches_the_oracle_table():
    import numpy as np
    from oracle.pyoracle import default_species_table as oracle_table
    from particlesim_b200 import default_species_table
    a, b = default_species_table(), oracle_table()
    assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert abs(float(a["lj_epsilon"][9]) - 9.94e-5) < 1e-6  # LJ_FORCE_EPSILON, config.rs:113


def test_no_cpu_fallback():
    """Without a GPU psim_create must fail with PSIM_E_CUDA; with one it must succeed."""
    from particlesim_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.psim_create(0, 16, 16, None, C.byref(h))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = None
    if has_gpu is False:
        assert rc == -1 and not h.value
    elif rc == 0:
        lib.psim_destroy(h)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "particlesim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may mention the test harness; code must not import or load the oracle
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text, f


def build_c_smoke():
    """tests/c/smoke.c -> tests/c/_build/smoke with plain gcc in C99 mode against include/psim_b200.h"""
    import subprocess
    from particlesim_b200 import _lib
    _lib.load()
    out_dir = os.path.join(ROOT, "tests", "c", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "smoke")
    libdir = os.path.join(ROOT, "particlesim_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "smoke.c"), "-o", exe, "-L", libdir, "-lpsim_b200",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return exe


def test_header_is_plain_c_and_a_c_client_links():
    """the boundary is a C ABI: a C99 translation unit includes the header, compiles warning-free and links
    against libpsim_b200.so (running it needs a GPU: tests/test_gpu_tree.py::test_c_client)"""
    exe = build_c_smoke()
    assert os.path.exists(exe)


def test_rust_binding_covers_the_header():
    """rust/psim-b200-sys: ffi.rs is generated from the header (every declared symbol bound, argument for argument),
    the #[repr(C)] structs list the header's fields in order, and the safe wrappers only call declared symbols"""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode == 0, \
        "rust/psim-b200-sys/src/ffi.rs is stale: run tools/gen_rust_ffi.py"
    ffi = open(os.path.join(ROOT, "rust", "psim-b200-sys", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (psim_\w+)\(", ffi))
    assert bound == set(declared_symbols())
    lib = open(os.path.join(ROOT, "rust", "psim-b200-sys", "src", "lib.rs")).read()
    used = set(re.findall(r"\b(psim_[a-z0-9_]+)\(", lib))
    assert used <= bound, used - bound
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "psim_b200.h")).read(), flags=re.S)
    for name in ("psim_config", "psim_species", "psim_node", "psim_stats", "psim_step_params"):
        body = re.search(r"typedef struct \{([^}]*)\} " + name + ";", header).group(1)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1]
            c_fields += [re.sub(r"\[.*", "", f.strip()) for f in names.split(",")]
        rust = re.search(r"pub struct " + name + r" \{(.*?)\n\}", lib, flags=re.S).group(1)
        r_fields = re.findall(r"pub (\w+):", rust)
        assert c_fields == r_fields, (name, c_fields, r_fields)
