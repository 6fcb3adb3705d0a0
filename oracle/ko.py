"""TEST INFRASTRUCTURE -- ctypes/numpy front-end to the plain-C oracle (oracle/kripke_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (kripke_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libkripke_oracle.so")
LAYOUTS = ["DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"]


class KoInput(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("num_directions", C.c_int), ("num_groups", C.c_int), ("legendre_order", C.c_int),
                ("quad_num_polar", C.c_int), ("quad_num_azimuthal", C.c_int), ("layout", C.c_int),
                ("npx", C.c_int), ("npy", C.c_int), ("npz", C.c_int),
                ("num_dirsets", C.c_int), ("num_groupsets", C.c_int), ("num_zonesets_dim", C.c_int * 3),
                ("sigt", C.c_double * 3), ("sigs", C.c_double * 3), ("num_material_subsamples", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "kripke_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.ko_create.restype = C.c_void_p
        L.ko_create.argtypes = [C.POINTER(KoInput)]
        for f in ("ko_destroy", "ko_ltimes", "ko_lplustimes", "ko_scattering", "ko_source"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.ko_zero.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_sweep_subdomain.argtypes = [C.c_void_p, C.c_int]
        L.ko_sweep_solver.argtypes = [C.c_void_p, C.c_int]
        L.ko_population.argtypes = [C.c_void_p]
        L.ko_population.restype = C.c_double
        L.ko_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ko_num_subdomains.argtypes = [C.c_void_p]
        L.ko_sweep_order.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ko_adjacency.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ko_field_chunks.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_field_chunk.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.ko_dim.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_default_input.argtypes = [C.POINTER(KoInput)]
        L.ko_check_input.argtypes = [C.POINTER(KoInput)]
    return _lib


def make_input(**kw):
    """kwargs mirror the kripke command line: zones=(x,y,z), groups, quad (int or (polar, azim)),
    legendre, layout ('DGZ'...), procs, dset, gset, zset, sigt, sigs."""
    inp = KoInput()
    lib().ko_default_input(C.byref(inp))
    if "zones" in kw:
        inp.nx, inp.ny, inp.nz = kw["zones"]
    if "groups" in kw:
        inp.num_groups = kw["groups"]
    if "quad" in kw:
        q = kw["quad"]
        if isinstance(q, (tuple, list)):
            inp.quad_num_polar, inp.quad_num_azimuthal = q
            inp.num_directions = q[0] * q[1]
        else:
            inp.num_directions = q
    if "legendre" in kw:
        inp.legendre_order = kw["legendre"]
    if "layout" in kw:
        inp.layout = LAYOUTS.index(kw["layout"].upper())
    if "procs" in kw:
        inp.npx, inp.npy, inp.npz = kw["procs"]
    if "dset" in kw:
        inp.num_dirsets = kw["dset"]
    if "gset" in kw:
        inp.num_groupsets = kw["gset"]
    if "zset" in kw:
        for i in range(3):
            inp.num_zonesets_dim[i] = kw["zset"][i]
    for name in ("sigt", "sigs"):
        if name in kw:
            for i in range(3):
                getattr(inp, name)[i] = kw[name][i]
    return inp


class Problem:
    def __init__(self, **kw):
        self.inp = make_input(**kw)
        self.h = lib().ko_create(C.byref(self.inp))
        if not self.h:
            raise ValueError("oracle: invalid input / unsupported quadrature")

    def close(self):
        if self.h:
            lib().ko_destroy(self.h)
            self.h = None

    __del__ = close

    def dim(self, which):
        return lib().ko_dim(self.h, which.encode())

    def num_chunks(self, name):
        return lib().ko_field_chunks(self.h, name.encode())

    def chunk(self, name, c):
        """numpy VIEW (no copy) of one chunk of a field."""
        ptr, n, es = C.c_void_p(), C.c_size_t(), C.c_int()
        rc = lib().ko_field_chunk(self.h, name.encode(), c, C.byref(ptr), C.byref(n), C.byref(es))
        if rc:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, dtype=np.float64 if es.value == 8 else np.int32)
        ctype = C.c_double if es.value == 8 else C.c_int
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n.value,))
        return arr

    def field(self, name):
        """concatenation (copy) of all chunks in work-list order, as the reference driver dumps them."""
        return np.concatenate([self.chunk(name, c) for c in range(self.num_chunks(name))])

    def zero(self, name):
        lib().ko_zero(self.h, name.encode())

    def ltimes(self):
        lib().ko_ltimes(self.h)

    def lplustimes(self):
        lib().ko_lplustimes(self.h)

    def scattering(self):
        lib().ko_scattering(self.h)

    def source(self):
        lib().ko_source(self.h)

    def sweep_subdomain(self, s):
        lib().ko_sweep_subdomain(self.h, s)

    def sweep_solver(self, bj=False):
        lib().ko_sweep_solver(self.h, int(bj))

    def population(self):
        return lib().ko_population(self.h)

    def solve(self, niter, bj=False):
        out = (C.c_double * niter)()
        lib().ko_solve(self.h, niter, int(bj), out)
        return list(out)

    def num_subdomains(self):
        return lib().ko_num_subdomains(self.h)

    def sweep_order(self):
        n = self.num_subdomains()
        out = (C.c_int * n)()
        lib().ko_sweep_order(self.h, out)
        return list(out)

    def adjacency(self, s):
        up, dn = (C.c_int * 3)(), (C.c_int * 3)()
        lib().ko_adjacency(self.h, s, up, dn)
        return list(up), list(dn)
