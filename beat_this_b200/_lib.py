"""ctypes binding of libbeatthis_sm100.so (C ABI in include/beatthis.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``beat_this_b200._lib.build()``
(nvcc, sm_100a).  There is no fallback: if the shared object is missing or no sm_100 GPU is
present, loading / ``bt_create`` fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# BT_LIB_PATH: an instrumented build of the same sources (e.g. build(extra_flags=("-DBT_FF_PROF",), out_path=...))
LIB_PATH = os.environ.get("BT_LIB_PATH") or os.path.join(HERE, "libbeatthis_sm100.so")
SOURCES = ["bt_api.cu", "kernels_simt.cu", "kernels_misc.cu", "kernels_gemm.cu", "kernels_attn.cu", "kernels_fused.cu", "dbn_host.cpp", "host_stage.cpp"]
HEADERS = ["common.cuh", "epilogue.cuh", "tc_common.cuh", "bt_kernels.h", os.path.join("..", "..", "include", "beatthis.h")]

BT_DTYPE_F32 = 0
BT_DTYPE_H16 = 1


class bt_wav_info(ctypes.Structure):
    _fields_ = [
        ("sample_rate", c_int32),
        ("channels", c_int32),
        ("bytes_per_sample", c_int32),
        ("is_float", c_int32),
        ("frames", c_int64),
        ("data_offset", c_int64),
    ]


class bt_hparams(ctypes.Structure):
    _fields_ = [
        ("spect_dim", c_int32),
        ("transformer_dim", c_int32),
        ("ff_mult", c_int32),
        ("n_layers", c_int32),
        ("head_dim", c_int32),
        ("stem_dim", c_int32),
        ("sum_head", c_int32),
        ("partial_transformers", c_int32),
    ]


# every symbol include/beatthis.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "bt_version": (c_int, []),
    "bt_act_dtype": (c_char_p, []),
    "bt_create": (c_int, [POINTER(c_void_p), c_int, POINTER(bt_hparams), c_int]),
    "bt_set_param": (c_int, [c_void_p, c_char_p, POINTER(c_float), c_int64]),
    "bt_finalize": (c_int, [c_void_p]),
    "bt_destroy": (None, [c_void_p]),
    "bt_last_error": (c_char_p, [c_void_p]),
    "bt_num_frames": (c_int64, [c_int64]),
    "bt_plan_chunks": (c_int64, [c_int64, POINTER(c_int64), POINTER(c_int64), c_int64]),
    "bt_logmel": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, POINTER(c_int64), c_void_p]),
    "bt_stage_audio": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32]),
    "bt_wav_probe": (c_int, [c_char_p, POINTER(bt_wav_info)]),
    "bt_stage_wav_files": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p]),
    "bt_resample": (
        c_int,
        [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, POINTER(c_int64), c_void_p],
    ),
    "bt_dbn_track": (
        c_int,
        [c_void_p, c_void_p, c_int32, c_void_p, c_int32, ctypes.c_double, ctypes.c_double, c_int32, ctypes.c_double,
         ctypes.c_double, ctypes.c_double, c_int32, ctypes.c_double, c_int32, c_void_p, c_void_p, c_void_p],
    ),
    "bt_dbn_viterbi": (
        c_int,
        [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "bt_spect2frames": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p, c_void_p]),
    "bt_forward_chunks": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "bt_audio2frames": (
        c_int,
        [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p, POINTER(c_int64), c_void_p],
    ),
    "bt_peakpick": (
        c_int,
        [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p],
    ),
    "bt_set_wave_chunks": (c_int, [c_void_p, c_int32]),
    "bt_launch_count": (c_int64, [c_void_p]),
    "bt_profile_enable": (c_int, [c_void_p, c_int]),
    "bt_profile_collect": (c_int, [c_void_p]),
    "bt_profile_reset": (c_int, [c_void_p]),
    "bt_profile_count": (c_int, [c_void_p]),
    "bt_profile_get": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_double), POINTER(c_int64)]),
    "bt_debug_request_tap": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "bt_debug_tap_count": (c_int64, [c_void_p]),
    "bt_debug_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "bt_debug_attention_time": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "bt_debug_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
}

_lib = None


OBJ_DIR = os.path.join(CSRC, "_obj")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def nvcc_command(out_path: str = LIB_PATH) -> list[str]:
    """The one-shot command line equivalent to what build() does (documentation / manual builds)."""
    return [_nvcc(), *NVCC_FLAGS, "-shared", *[os.path.join(CSRC, s) for s in SOURCES], "-o", out_path]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: tuple = (), out_path: str = LIB_PATH) -> str:
    """Compile the CUDA library for sm_100a (cross-compiles without a GPU): one nvcc -c per source, in parallel,
    objects cached under csrc/_obj (keyed by flags), then one link step."""
    from concurrent.futures import ThreadPoolExecutor

    if not (force or out_path != LIB_PATH or needs_build()):
        return LIB_PATH
    tag = "".join(c if c.isalnum() else "_" for c in "".join(extra_flags)) or "default"
    odir = os.path.join(OBJ_DIR, tag)
    os.makedirs(odir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)

    def compile_one(src):
        sp = os.path.join(CSRC, src)
        obj = os.path.join(odir, os.path.splitext(src)[0] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hdr_t):
            return obj, None
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", sp, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, (None if res.returncode == 0 else f"{src}:\n{res.stdout}\n{res.stderr}")

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    errors = [e for _, e in results if e]
    if errors:
        raise RuntimeError("nvcc failed:\n" + "\n".join(errors))
    tmp = out_path + f".tmp{os.getpid()}"
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *[o for o, _ in results], "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, out_path)
    return out_path


def load() -> ctypes.CDLL:
    """Load the shared library and declare all prototypes.  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or beat_this_b200._lib.build()). "
            "There is no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class BTError(RuntimeError):
    pass


def check(lib, ctx, code: int):
    if code != 0:
        msg = lib.bt_last_error(ctx)
        raise BTError(f"libbeatthis error {code}: {msg.decode() if msg else ''}")


def i64_array(values):
    arr = (c_int64 * len(values))(*[int(v) for v in values])
    return arr
