"""ctypes binding of libtdt_tab.so (include/tdt_tab.h): the signal tab files as numpy columns + interned string tables.

`TabSet` owns one native record set: every file parsed into it shares the interning context (read names, contig names,
orientation strings get ids in order of first appearance).  Columns are copied out as numpy arrays; the read-name table
stays in native memory behind `BlobStrings`, a read-only sequence that decodes a name only when it is asked for (the
cluster stage looks at the names of candidate members only)."""
import ctypes
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtdt_tab.so")
_vp, _i64 = ctypes.c_void_p, ctypes.c_int64
TDT_TAB_IRREGULAR = -4
KIND = {"discordants": 0, "splits": 1, "contigs": 2}
TAB_SIGNATURES = {
    "tdt_tab_last_error": (ctypes.c_char_p, []),
    "tdt_tab_new": (_vp, []),
    "tdt_tab_free": (None, [_vp]),
    "tdt_tab_parse": (_i64, [_vp, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "tdt_tab_n": (_i64, [_vp]),
    "tdt_tab_col_i32": (_vp, [_vp, ctypes.c_int]),
    "tdt_tab_col_i64": (_vp, [_vp, ctypes.c_int]),
    "tdt_tab_table": (_i64, [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
}
_lib = None


def tab_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m tiddit_b200.build`" % _LIB_PATH)
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in TAB_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class IrregularTab(Exception):
    """The file is not perfectly regular; read it line by line like the reference does."""


class BlobStrings:
    """strings[i] of a blob + offsets table, decoded on demand (read-only sequence)."""

    def __init__(self, blob, offsets):
        self._blob, self._off = blob, offsets

    def __len__(self):
        return len(self._off) - 1

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        return self._blob[self._off[i]:self._off[i + 1]].decode("utf-8")

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def tolist(self):
        return list(self)


class TabSet:
    def __init__(self):
        self.L = tab_lib()
        self.h = self.L.tdt_tab_new()

    def close(self):
        if self.h:
            self.L.tdt_tab_free(self.h)
            self.h = None

    __del__ = close

    def parse(self, path, kind, threads=0):
        """-> (first record, one past the last record) of the file inside the set; raises IrregularTab / OSError."""
        lo = int(self.L.tdt_tab_n(self.h))
        rc = int(self.L.tdt_tab_parse(self.h, os.fsencode(path), KIND[kind] if isinstance(kind, str) else int(kind), int(threads)))
        if rc == TDT_TAB_IRREGULAR:
            raise IrregularTab(self.L.tdt_tab_last_error().decode("utf-8", "replace"))
        if rc < 0:
            raise OSError(self.L.tdt_tab_last_error().decode("utf-8", "replace"))
        return lo, lo + rc

    def __len__(self):
        return int(self.L.tdt_tab_n(self.h))

    def _col(self, getter, which, dtype):
        n = len(self)
        if n == 0:
            return np.zeros(0, dtype=dtype)
        ptr = getter(self.h, which)
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()

    def col_i32(self, which):
        return self._col(self.L.tdt_tab_col_i32, which, np.int32)

    def col_i64(self, which):
        return self._col(self.L.tdt_tab_col_i64, which, np.int64)

    def table(self, which, lazy=False):
        """String table `which` (0 names, 1 contigs, 2 orientations): a list, or a BlobStrings when lazy."""
        blob, offs = _vp(), _vp()
        k = int(self.L.tdt_tab_table(self.h, which, ctypes.byref(blob), ctypes.byref(offs)))
        if k == 0:
            return BlobStrings(b"", np.zeros(1, dtype=np.int64)) if lazy else []
        off = np.ctypeslib.as_array(ctypes.cast(offs, ctypes.POINTER(ctypes.c_int64)), shape=(k + 1,)).copy()
        data = ctypes.string_at(blob, int(off[-1]))
        strings = BlobStrings(data, off)
        return strings if lazy else list(strings)
