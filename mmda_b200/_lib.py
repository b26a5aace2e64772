"""ctypes binding of libmmda_b200.so.  Prototypes are parsed from include/mmda_b200.h so the
header stays the single source of truth for the C ABI."""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "mmda_b200.h")
LIB_PATH = os.path.join(HERE, "libmmda_b200.so")

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong, "unsigned": ctypes.c_uint,
    "mmda_stream_t": ctypes.c_void_p,
}
_RET = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "const char*": ctypes.c_char_p}


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(mmda_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        args = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    args.append(ctypes.c_void_p)
                else:
                    ty = p.rsplit(" ", 1)[0]
                    args.append(_SCALARS[ty])
        protos[name] = (_RET[ret], args)
    return protos


class MmdaError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise MmdaError(
                    f"{LIB_PATH} is missing: build it with `python -m mmda_b200.build` "
                    "(there is no CPU or PyTorch fallback for the MISA hot path)")
            dll = ctypes.CDLL(LIB_PATH)
            for name, (ret, args) in self.protos.items():
                fn = getattr(dll, name)
                fn.restype, fn.argtypes = ret, args
            self._dll = dll
        return self._dll

    def raw(self, name):
        return getattr(self.load(), name)

    def call(self, name, *args):
        rc = getattr(self.load(), name)(*args)
        if rc != 0:
            msg = self._dll.mmda_last_error().decode(errors="replace")
            raise MmdaError(f"{name} failed (rc={rc}): {msg}")
        return rc


LIB = _Lib()
