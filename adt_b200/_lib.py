"""ctypes binding of libadt_b200.so.

The ctypes.Structure classes are generated from include/adt_b200.h at import time, so the Python side can
never drift from the C ABI.  There is NO fallback: if the shared library is missing the import of any compute
entry point raises (the product path must fail loudly, not silently run on the CPU).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "adt_b200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libadt_b200.so")

_SCALARS = {
    "int32_t": ctypes.c_int32, "uint32_t": ctypes.c_uint32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "double": ctypes.c_double, "int": ctypes.c_int,
}


def _parse_header(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"^(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl)
            base, star, names = m.group(2), m.group(3), m.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                ptr = bool(star)
                if nm.startswith("*"):
                    ptr, nm = True, nm[1:].strip()
                if ptr:
                    ctype = ctypes.c_void_p
                elif base in _SCALARS:
                    ctype = _SCALARS[base]
                else:
                    ctype = structs[base]
                fields.append((nm, ctype))
        structs[name] = type(name, (ctypes.Structure,), {"_fields_": fields})
    funcs = re.findall(r"^\s*(?:int|const char\*)\s+(adt_\w+)\s*\(", src, flags=re.M)
    return structs, sorted(set(funcs))


STRUCTS, FUNCTIONS = _parse_header(HEADER)
globals().update(STRUCTS)

_lib = None


class AdtError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AdtError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the ADT hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.adt_last_error.restype = ctypes.c_char_p
        for f in FUNCTIONS:
            if f != "adt_last_error":
                getattr(_lib, f).restype = ctypes.c_int
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise AdtError(f"{what} failed ({rc}): {lib().adt_last_error().decode()}")


def ptr(t):
    """device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def fill(struct, **kw):
    """populate a ctypes struct; torch tensors become device pointers, None becomes NULL."""
    for k, v in kw.items():
        if hasattr(v, "data_ptr"):
            v = v.data_ptr()
        setattr(struct, k, v)
    return struct
