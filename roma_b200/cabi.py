"""ctypes binding of libromab200.so — the thin shim between PyTorch (device memory, streams) and the C ABI.

The argument structs are generated from `include/romab200.h` itself at import time, so the Python side
cannot drift from the header.  Every wrapper passes raw device pointers (`tensor.data_ptr()`) and the
current CUDA stream; a non-zero return code becomes a `RuntimeError` carrying `romab200_last_error()`.
There is no fallback: if the library is missing or the device is not a B200, calls fail loudly.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "romab200.h")
LIB_PATH = os.path.join(HERE, "lib", "libromab200.so")

RB_F32, RB_F16, RB_BF16, RB_F16S = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ROWMAP_NONE, ROWMAP_PAD_KEEP, ROWMAP_PAD_TO_COMPACT, ROWMAP_SEGMENT = 0, 1, 2, 3
EPI_LINEAR, EPI_COSKERNEL = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2

DTYPE_CODE = {torch.float32: RB_F32, torch.float16: RB_F16, torch.bfloat16: RB_BF16}

_CTYPES = {
    "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
}


def _parse_header(path: str):
    """Returns ({struct_name: [(field, ctype)]}, [function names]) parsed from the C header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    structs: Dict[str, list] = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl, flags=re.S)
            base, ptr, names = m.group(2), m.group(3), m.group(4)
            first = True
            for part in names.split(","):
                part = part.strip()
                is_ptr = bool(ptr) if first else part.startswith("*")
                first = False
                part = part.lstrip("*").strip()
                arr = re.match(r"(\w+)\[(\d+)\]$", part)
                if is_ptr:
                    fields.append((part, ctypes.c_void_p))
                elif arr:
                    fields.append((arr.group(1), _CTYPES[base] * int(arr.group(2))))
                else:
                    fields.append((part, _CTYPES[base]))
        structs[name] = fields
    funcs = re.findall(r"\b(romab200_\w+)\s*\(", text)
    return structs, sorted(set(funcs))


STRUCT_FIELDS, FUNCTIONS = _parse_header(HEADER)
STRUCTS = {name: type(name, (ctypes.Structure,), {"_fields_": fields}) for name, fields in STRUCT_FIELDS.items()}

_lib = None


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    """dlopen the in-tree library and type its entry points.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m roma_b200.build` "
                           "(there is no CPU or PyTorch fallback for the CUDA path)")
    lib = ctypes.CDLL(path)
    lib.romab200_last_error.restype = ctypes.c_char_p
    lib.romab200_abi_version.restype = ctypes.c_int
    lib.romab200_device_ok.restype = ctypes.c_int
    lib.romab200_launch_count.restype = ctypes.c_ulonglong
    for fn in FUNCTIONS:
        getattr(lib, fn)            # AttributeError if the header declares a symbol the library lacks
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


launch_count = 0      # number of C-ABI calls made (bench.py reports kernel launches from it)


def call(fn_name: str, struct_name: str, **kw) -> None:
    """Fill `struct_name` from keyword arguments (tensors become device pointers) and call `fn_name`."""
    global launch_count
    lib = load_library()
    args = STRUCTS[struct_name]()
    valid = {f for f, _ in STRUCT_FIELDS[struct_name]}
    for k, v in kw.items():
        if k not in valid:
            raise TypeError(f"{struct_name} has no field {k}")
        if isinstance(v, torch.Tensor) or v is None:
            setattr(args, k, _ptr(v))
        elif isinstance(v, (list, tuple)):
            arr = getattr(args, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(args, k, v)
    rc = getattr(lib, fn_name)(ctypes.byref(args), _stream())
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{fn_name} failed: {lib.romab200_last_error().decode()}")


def kernel_launches() -> int:
    """Kernels launched by libromab200 in this process so far."""
    return int(load_library().romab200_launch_count())


def device_ok() -> bool:
    return bool(load_library().romab200_device_ok())
