"""ctypes bindings of the two native libraries (no compute here)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LAYOUTS = ["DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"]
KB200_MAX_DIRSETS = 64


class KB200Error(RuntimeError):
    pass


def lib_paths():
    return (os.path.join(_HERE, "lib", "libkripke_b200.so"), os.path.join(_HERE, "lib", "libkripke_host.so"))


def build(verbose=False):
    """Compile the CUDA library (nvcc, sm_100a), the host library and kripke.exe in-tree."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j8", "all"], stdout=out)
    return lib_paths()


_abi = None
_host = None


def abi():
    """libkripke_b200.so -- the C ABI declared in include/kripke_b200.h."""
    global _abi
    if _abi is None:
        path = lib_paths()[0]
        if not os.path.exists(path):
            raise KB200Error(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.kb200_last_error.restype = C.c_char_p
        L.kb200_init.argtypes = [C.c_int]
        L.kb200_device_count.argtypes = [C.POINTER(C.c_int)]
        L.kb200_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.kb200_free.argtypes = [C.c_void_p]
        L.kb200_alloc_host.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.kb200_free_host.argtypes = [C.c_void_p]
        L.kb200_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.kb200_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.kb200_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.kb200_fill_f64.argtypes = [C.c_void_p, C.c_double, C.c_size_t, C.c_void_p]
        L.kb200_stream_sync.argtypes = [C.c_void_p]
        L.kb200_stream_create.argtypes = [C.POINTER(C.c_void_p)]
        L.kb200_stream_destroy.argtypes = [C.c_void_p]
        L.kb200_stream_wait_event.argtypes = [C.c_void_p, C.c_void_p]
        L.kb200_event_create.argtypes = [C.POINTER(C.c_void_p)]
        L.kb200_event_destroy.argtypes = [C.c_void_p]
        L.kb200_event_record.argtypes = [C.c_void_p, C.c_void_p]
        L.kb200_event_sync.argtypes = [C.c_void_p]
        L.kb200_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.kb200_launch_count.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.kb200_population_scratch_doubles.restype = C.c_size_t
        L.kb200_layout_transform.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kb200_comm_unique_id.argtypes = [C.c_void_p]
        L.kb200_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.kb200_peak_fp64_gflops.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.kb200_peak_copy_gbs.argtypes = [C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.kb200_device_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        _abi = L
    return _abi


def check(rc, what="kb200 call"):
    if rc:
        raise KB200Error(f"{what} failed ({rc}): {abi().kb200_last_error().decode()}")


def host():
    """libkripke_host.so -- flat C entry points over the C++ Kripke:: host layer (host/capi.cpp)."""
    global _host
    if _host is None:
        abi()
        path = lib_paths()[1]
        if not os.path.exists(path):
            raise KB200Error(f"{path} is missing: run __graft_entry__.build()")
        H = C.CDLL(path, mode=C.RTLD_GLOBAL)
        H.kripke_b200_create.restype = C.c_void_p
        H.kripke_b200_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        H.kripke_b200_destroy.argtypes = [C.c_void_p]
        H.kripke_b200_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        H.kripke_b200_call.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        H.kripke_b200_field_chunks.argtypes = [C.c_void_p, C.c_char_p]
        H.kripke_b200_field_chunk_size.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        H.kripke_b200_field_chunk_size.restype = C.c_long
        H.kripke_b200_field_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        H.kripke_b200_field_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        H.kripke_b200_field_device_ptr.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        H.kripke_b200_field_device_ptr.restype = C.c_void_p
        H.kripke_b200_release_host_mirrors.argtypes = [C.c_void_p, C.c_char_p]
        H.kripke_b200_timer_total.argtypes = [C.c_void_p, C.c_char_p]
        H.kripke_b200_timer_total.restype = C.c_double
        H.kripke_b200_timer_count.argtypes = [C.c_void_p, C.c_char_p]
        H.kripke_b200_timer_count.restype = C.c_long
        H.kripke_b200_num_unknowns.argtypes = [C.c_void_p]
        H.kripke_b200_num_unknowns.restype = C.c_long
        H.kripke_b200_num_subdomains.argtypes = [C.c_void_p]
        H.kripke_b200_niter.argtypes = [C.c_void_p]
        H.kripke_b200_is_bj.argtypes = [C.c_void_p]
        H.kripke_b200_set_world.argtypes = [C.c_int, C.c_int]
        H.kripke_b200_timer_sync.argtypes = [C.c_int]
        H.kripke_b200_sweep_schedule.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int)]
        _host = H
    return _host


def have_gpu():
    n = C.c_int(0)
    try:
        return abi().kb200_device_count(C.byref(n)) == 0 and n.value > 0
    except KB200Error:
        return False


_device = None


def init_device(device=0):
    """Bind this process to one B200.  Raises when no sm_100 GPU is visible (no CPU fallback)."""
    global _device
    if _device != device:
        check(abi().kb200_init(device), "kb200_init")
        _device = device
    return device


class Problem:
    """A generated Kripke problem (the reference's DataStore) driven through the C++ host layer."""

    def __init__(self, args, quiet=True):
        if isinstance(args, str):
            args = args.split()
        argv = [b"kripke"] + [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        saved = None
        if quiet:  # the host layer prints the reference's banner blocks with printf
            import sys
            sys.stdout.flush()
            saved = os.dup(1)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            os.close(devnull)
        try:
            self.h = host().kripke_b200_create(len(argv), arr)
        finally:
            if saved is not None:
                C.CDLL(None).fflush(None)
                os.dup2(saved, 1)
                os.close(saved)
        if not self.h:
            raise ValueError("invalid kripke command line: " + " ".join(map(str, args)))
        self.args = list(args)
        self.quiet = quiet

    def close(self):
        if getattr(self, "h", None):
            host().kripke_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _quiet_call(self, fn, *a):
        if not self.quiet:
            return fn(*a)
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        os.close(devnull)
        try:
            return fn(*a)
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)

    # --- solver / entry points -----------------------------------------------------------
    def solve(self, niter=None):
        """Kripke::SteadyStateSolver; returns the per-iteration particle counts (full precision)."""
        n = niter if niter is not None else host().kripke_b200_niter(self.h)
        out = (C.c_double * n)()
        self._quiet_call(host().kripke_b200_solve, self.h, n, out)
        return list(out)

    def call(self, what):
        """One reference entry point by name: LTimes, LPlusTimes, scattering, source, SweepSolver,
        population, sweepSubdomain:<id>, zero:<field>."""
        res = C.c_double(0.0)
        rc = self._quiet_call(host().kripke_b200_call, self.h, what.encode(), C.byref(res))
        if rc:
            raise KeyError(what)
        return res.value

    # --- fields ----------------------------------------------------------------------------
    def num_chunks(self, name):
        n = host().kripke_b200_field_chunks(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        return n

    def chunk(self, name, c):
        es = C.c_int(0)
        n = host().kripke_b200_field_chunk_size(self.h, name.encode(), c, C.byref(es))
        dt = {8: np.float64, 4: np.int32}[es.value] if name not in ("upwind", "downwind", "SdomId2GlobalSdomId",
                                                                   "GlobalSdomId2Rank", "GlobalSdomId2SdomId") else np.int64
        out = np.empty(n, dtype=dt)
        host().kripke_b200_field_get(self.h, name.encode(), c, out.ctypes.data_as(C.c_void_p))
        return out

    def set_chunk(self, name, c, values):
        cur = self.chunk(name, c)
        v = np.ascontiguousarray(values, dtype=cur.dtype)
        assert v.shape == cur.shape
        host().kripke_b200_field_set(self.h, name.encode(), c, v.ctypes.data_as(C.c_void_p))

    def field(self, name):
        return np.concatenate([self.chunk(name, c) for c in range(self.num_chunks(name))])

    def norm2(self, name, release=True):
        """||field||_2 accumulated chunk by chunk (blocked dot products summed with math.fsum), so that multi-GB fields
        never need more host memory than one chunk; the chunk's host mirror is released after use."""
        import math
        parts = []
        n = 0
        for c in range(self.num_chunks(name)):
            v = self.chunk(name, c)
            n += len(v)
            for i in range(0, len(v), 1 << 20):
                b = v[i:i + (1 << 20)]
                parts.append(float(np.dot(b, b)))
            del v
            if release:
                self.release_host_mirrors(name)
        return math.sqrt(math.fsum(parts)), n

    def device_ptr(self, name, c, will_write=False):
        return host().kripke_b200_field_device_ptr(self.h, name.encode(), c, int(will_write))

    def release_host_mirrors(self, name):
        host().kripke_b200_release_host_mirrors(self.h, name.encode())

    # --- misc ------------------------------------------------------------------------------
    def timer(self, name):
        return host().kripke_b200_timer_total(self.h, name.encode())

    def timer_count(self, name):
        return host().kripke_b200_timer_count(self.h, name.encode())

    def num_unknowns(self):
        return host().kripke_b200_num_unknowns(self.h)

    def num_subdomains(self):
        return host().kripke_b200_num_subdomains(self.h)

    def sweep_schedule(self):
        n = self.num_subdomains()
        order, stage = (C.c_int * n)(), (C.c_int * n)()
        rf, st = (C.c_int * (3 * n))(), (C.c_int * (3 * n))()
        k = host().kripke_b200_sweep_schedule(self.h, order, stage, rf, st)
        return [dict(sdom=order[i], stage=stage[i], recv_from=list(rf[3 * i:3 * i + 3]), send_to=list(st[3 * i:3 * i + 3]))
                for i in range(k)]
