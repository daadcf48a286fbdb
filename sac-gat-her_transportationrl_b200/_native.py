"""Loader and builder of the C-ABI CUDA library (``include/b200_assign.h``).

The library is built IN-TREE by ``nvcc`` for sm_100a only and loaded with ``ctypes``; there is no
CPU fallback - every compute entry point raises if the library (or a CUDA device) is missing.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess


PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("B200_LIB", os.path.join(PKG_DIR, "libb200assign.so"))  # env override: kernel A/B runs
SOURCES = ["b200_graph.cu", "b200_cch.cu", "b200_assign.cu", "b200_assign_small.cu", "b200_assign_static.cu", "b200_env.cu",
           "b200_gat.cu"]
HEADERS = ["b200_common.cuh", "b200_assign_internal.cuh", "b200_assign_small.cuh", "b200_assign_static.cuh", "b200_assign_static_grp.cuh",
           "b200_topo_siouxfalls.h", os.path.join("..", "..", "include", "b200_assign.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # link costs / flow updates are separately rounded IEEE ops (parity with the oracle)
    "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/b200_assign.h declares (tests/test_host.py::test_c_abi_library_loads_and_exports_header_symbols)
EXPORTS = [
    "b200_last_error", "b200_version", "b200_graph_create", "b200_graph_destroy", "b200_graph_info",
    "b200_graph_origins", "b200_bpr", "b200_sssp", "b200_aon", "b200_update_reduce", "b200_assign",
    "b200_assign_launches", "b200_assign_last_config", "b200_cch_host_sssp", "b200_assign_kernel_name", "b200_register_static_plugin", "b200_state_features", "b200_apply_actions", "b200_sample_damage", "b200_reward_done",
    "b200_gat_edge_softmax_aggregate_fwd", "b200_gat_edge_softmax_aggregate_bwd",
]


class NativeLibraryError(RuntimeError):
    pass


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise NativeLibraryError("nvcc not found; the CUDA extension cannot be built")
    return exe


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def _obj_dir() -> str:
    """csrc/_obj for the product build; a tuning build (B200_NVCC_EXTRA) keeps its objects apart so that they can
    never be linked into the product library by a later incremental build"""
    extra = os.environ.get("B200_NVCC_EXTRA", "").strip()
    return os.path.join(CSRC, "_obj" if not extra else "_obj_" + hashlib.sha1(extra.encode()).hexdigest()[:8])


def _obj_path(src: str) -> str:
    return os.path.join(_obj_dir(), os.path.splitext(src)[0] + ".o")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libb200assign.so (sm_100a).  nvcc cross-compiles without a GPU.
    One object per source, compiled in parallel, then linked; objects live in csrc/_obj/."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(_obj_dir(), exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    compile_flags += os.environ.get("B200_NVCC_EXTRA", "").split()  # kernel-tuning experiments (-DB200_STATIC_MINB=..)
    api = os.path.join("..", "..", "include", "b200_assign.h")
    deps = {"b200_assign.cu": ["b200_common.cuh", "b200_assign_internal.cuh", api],
            "b200_assign_small.cu": ["b200_common.cuh", "b200_assign_internal.cuh", "b200_assign_small.cuh", api],
            "b200_assign_static.cu": ["b200_common.cuh", "b200_assign_internal.cuh", "b200_assign_static.cuh",
                                      "b200_assign_static_grp.cuh", "b200_topo_siouxfalls.h", api],
            "b200_graph.cu": ["b200_common.cuh", api], "b200_env.cu": ["b200_common.cuh", api],
            "b200_cch.cu": ["b200_common.cuh", "b200_assign_internal.cuh", api],
            "b200_gat.cu": ["b200_common.cuh", api]}
    procs = []
    for src in SOURCES:
        obj, path = _obj_path(src), os.path.join(CSRC, src)
        hdr_time = max(os.path.getmtime(os.path.normpath(os.path.join(CSRC, h)))
                       for h in deps.get(src, ["b200_common.cuh", api]))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
            continue
        procs.append((src, subprocess.Popen([_nvcc()] + compile_flags + ["-c", "-o", obj, path],
                                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    log = []
    for src, proc in procs:
        _out, err = proc.communicate()
        log.append(err)
        if proc.returncode != 0:
            raise NativeLibraryError(f"nvcc failed on {src}:\n" + err[-4000:])
    res = subprocess.run([_nvcc(), "-shared", "-o", LIB_PATH] + [_obj_path(s) for s in SOURCES] + ["-ldl"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeLibraryError("link failed:\n" + res.stderr[-4000:])
    if verbose:
        print("\n".join(log))
    return LIB_PATH


_lib = None

_vp, _i, _d, _f = C.c_void_p, C.c_int, C.c_double, C.c_float


def _declare(lib) -> None:
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_assign_kernel_name.restype = C.c_char_p
    lib.b200_version.restype = _i
    lib.b200_graph_create.argtypes = [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_vp)]
    lib.b200_graph_destroy.argtypes = [_vp]
    lib.b200_graph_info.argtypes = [_vp, _vp]
    lib.b200_graph_origins.argtypes = [_vp, _vp]
    lib.b200_bpr.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _d, _d, _vp]
    lib.b200_sssp.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]
    lib.b200_aon.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp]
    lib.b200_update_reduce.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d,
                                       _i, _vp]
    lib.b200_assign.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _i, _vp]
    lib.b200_assign_launches.argtypes = [_vp]
    lib.b200_assign_last_config.argtypes = [_vp, _vp]
    lib.b200_cch_host_sssp.argtypes = [_i, _i, _vp, _vp, _vp, _i, _vp, _vp]
    lib.b200_assign_kernel_name.argtypes = [_vp]
    lib.b200_register_static_plugin.argtypes = [C.c_char_p]
    lib.b200_state_features.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _d, _vp, _vp, _vp, _vp]
    lib.b200_apply_actions.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp]
    lib.b200_sample_damage.argtypes = [_vp, _i, _i, _i, C.c_uint64, _vp, _vp, _vp, _vp, _vp]
    lib.b200_reward_done.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp, _vp, _vp]
    lib.b200_gat_edge_softmax_aggregate_fwd.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]
    lib.b200_gat_edge_softmax_aggregate_bwd.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp,
                                                        _vp, _vp, _vp]
    for name in EXPORTS:
        if name not in ("b200_last_error", "b200_assign_kernel_name"):
            getattr(lib, name).restype = _i


def lib():
    """The loaded library; raises NativeLibraryError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the assignment path)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


# ---------------------------------------------------------------------------------------------
# on-demand network-specialised kernels
# ---------------------------------------------------------------------------------------------
SPECIALIZE_MAX_NODES, SPECIALIZE_MAX_LINKS = 32, 160  # distance labels live in registers: 2 per node


def specialize(num_nodes: int, row, col, t0, origins, verbose: bool = False) -> str:
    """Compile (once, cached in csrc/_jit/) and register the straight-line assignment kernel for one
    small topology: the same hand-written template Sioux Falls uses (csrc/b200_assign_static.cuh),
    instantiated for a generated topology header.  Networks uploaded afterwards with this link list
    dispatch to it.  Takes about a minute of nvcc the first time; returns the plugin path."""
    import hashlib

    import numpy as np

    from . import codegen

    row = np.ascontiguousarray(row, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    E = int(row.shape[0])
    if num_nodes > SPECIALIZE_MAX_NODES or E > SPECIALIZE_MAX_LINKS or len(origins) > 32 or len(origins) < 1:
        raise ValueError(f"network too large for a specialised kernel (N <= {SPECIALIZE_MAX_NODES}, "
                         f"E <= {SPECIALIZE_MAX_LINKS}, origins <= 32)")
    h = hashlib.sha1()
    h.update(np.int64(num_nodes).tobytes() + row.tobytes() + col.tobytes())
    for dep in ("b200_static_plugin.cu", "b200_assign_static.cuh", "b200_assign_static_grp.cuh", "b200_assign_internal.cuh",
                "b200_common.cuh"):
        with open(os.path.join(CSRC, dep), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    key = h.hexdigest()[:16]
    jit_dir = os.path.join(CSRC, "_jit")
    so = os.path.join(jit_dir, f"topo_{key}.so")
    if not os.path.exists(so):
        os.makedirs(jit_dir, exist_ok=True)
        name = f"Topo_{key}"
        header = os.path.join(jit_dir, f"topo_{key}.h")
        with open(header, "w") as fh:
            fh.write(codegen.emit(name, name.upper(), num_nodes, row, col, np.asarray(t0, dtype=np.float32),
                                  [int(o) for o in origins], f"a {num_nodes}-node / {E}-link network"))
        cmd = [_nvcc()] + NVCC_FLAGS + [f'-DB200_TOPO_HEADER="{header}"', f"-DB200_TOPO_NAME={name}", "-o", so + ".tmp",
                                        os.path.join(CSRC, "b200_static_plugin.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise NativeLibraryError("nvcc failed on the specialised kernel:\n" + res.stderr[-4000:])
        os.replace(so + ".tmp", so)
        if verbose:
            print(res.stderr)
    check(lib().b200_register_static_plugin(so.encode()), "b200_register_static_plugin")
    return so


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = lib().b200_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise NativeLibraryError(f"{what} failed (code {rc}): {msg}")
