"""Host-side mirror of the reference's solver interface over the C-ABI.

`FBstabDense` / `FBstabMpc` keep the reference's names, argument meaning and
error behaviour (reference fbstab/fbstab_dense.h:50-194, fbstab_mpc.h:56-243):
construct with the problem sizes, `update_options`, `default_options`,
`reliable_options`, `solve` for one instance, plus the batched `solve_batch`.
Size errors raise RuntimeError like the reference's std::runtime_error.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import OUT_DTYPE, ComponentIO, Options


def _is_host(a):
    return isinstance(a, np.ndarray)


def _flat(a, n, name, device=None, dtype_bytes=8):
    """Checks one C-ABI argument before its raw pointer is handed over: element
    count, float64 (or `dtype_bytes`-wide records), contiguity and -- for CUDA
    tensors -- the device of the handle.  Raises RuntimeError (never assert: the
    kernels would read past a mis-typed allocation)."""
    if a is None:
        if n == 0:
            return None
        raise RuntimeError(f"{name} is required")
    if isinstance(a, np.ndarray):
        if a.size != n:
            raise RuntimeError(f"size mismatch for {name}: expected {n} elements, got {a.size}")
        if a.dtype.itemsize != dtype_bytes or (dtype_bytes == 8 and a.dtype != np.float64) or \
                not a.flags["C_CONTIGUOUS"]:
            raise RuntimeError(f"{name} must be contiguous float64")
        return a
    if not hasattr(a, "data_ptr"):
        raise RuntimeError(f"{name}: expected a numpy array or a torch tensor, got {type(a)}")
    if a.numel() * a.element_size() != n * dtype_bytes:
        raise RuntimeError(f"size mismatch for {name}: expected {n * dtype_bytes} bytes, "
                           f"got {a.numel() * a.element_size()}")
    if dtype_bytes == 8 and str(a.dtype) != "torch.float64":
        raise RuntimeError(f"{name} must be a float64 tensor, got {a.dtype}")
    if not a.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if a.is_cuda and device is not None and a.device.index != device:
        raise RuntimeError(f"{name} lives on cuda:{a.device.index}, the solver on cuda:{device}")
    return a


class _Base:
    _prefix = None
    _fields = ()

    def __init__(self):
        self._h = C.c_void_p()
        self.opts = capi.default_options()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                getattr(capi.lib(), f"fbstab_{self._prefix}_batch_destroy")(h)
            except TypeError:  # interpreter shutdown: the module globals are gone
                pass
            self._h = None

    # -- options: UpdateOptions / DefaultOptions / ReliableOptions ----------
    def update_options(self, opts):
        capi.check(getattr(capi.lib(), f"fbstab_{self._prefix}_batch_set_options")(
            self._h, C.byref(opts)))
        got = Options()
        capi.check(getattr(capi.lib(), f"fbstab_{self._prefix}_batch_get_options")(
            self._h, C.byref(got)))
        self.opts = got

    @staticmethod
    def default_options(**kw):
        return capi.default_options(**kw)

    @staticmethod
    def reliable_options(**kw):
        return capi.reliable_options(**kw)

    @property
    def path(self):
        return getattr(capi.lib(), f"fbstab_{self._prefix}_batch_path")(self._h).decode()

    @property
    def last_launches(self):
        return getattr(capi.lib(), f"fbstab_{self._prefix}_batch_last_launches")(self._h)

    def _data_ptrs(self, data, batch):
        sizes = self.field_sizes
        return [capi.ptr(_flat(data[k], batch * sizes[k], k, self.device)) for k in self._fields]

    def solve_batch(self, data, z, l, v, y=None, out=None, stream=None):
        """Solve `batch` instances.  `data` maps field name -> flat array
        (numpy host array or torch tensor, host or CUDA).  z,l,v are the warm
        start and are overwritten; returns (out, y)."""
        batch = (z.size if _is_host(z) else z.numel()) // self.nz
        if batch > self.max_batch:
            raise RuntimeError("batch exceeds max_batch")
        _flat(z, batch * self.nz, "z", self.device)
        _flat(l, batch * self.nl, "l", self.device)
        _flat(v, batch * self.nv, "v", self.device)
        if y is not None:
            _flat(y, batch * self.nv, "y", self.device)
        if out is not None:
            _flat(out, batch, "out", self.device, dtype_bytes=OUT_DTYPE.itemsize)
        if y is None:
            y = np.zeros(batch * self.nv) if _is_host(z) else z.new_zeros(batch * self.nv)
        if out is None:
            if _is_host(z):
                out = np.zeros(batch, dtype=OUT_DTYPE)
            else:
                import torch
                out = torch.zeros(batch * OUT_DTYPE.itemsize, dtype=torch.uint8,
                                  device=z.device)
        fn = getattr(capi.lib(), f"fbstab_{self._prefix}_batch_solve")
        capi.check(fn(self._h, batch, *self._data_ptrs(data, batch), capi.ptr(z),
                      capi.ptr(l), capi.ptr(v), capi.ptr(y), capi.ptr(out),
                      stream))
        return out, y

    def component(self, comp, data, batch, stream=None, **io):
        """Run one engine stage (see FBSTAB_COMP_* in the header)."""
        cio = ComponentIO()
        keep = []
        for name, _ in ComponentIO._fields_:
            if name in ("sigma", "tol"):
                setattr(cio, name, float(io.get(name, 0.0)))
            else:
                a = io.get(name)
                keep.append(a)
                setattr(cio, name, capi.ptr(a))
        fn = getattr(capi.lib(), f"fbstab_{self._prefix}_batch_component")
        capi.check(fn(self._h, comp, batch, *self._data_ptrs(data, batch),
                      C.byref(cio), stream))


class FBstabDense(_Base):
    """min 1/2 z'Hz + f'z  s.t. Gz = h, Az <= b   (fbstab_dense.h:17-49)."""
    _prefix = "dense"
    _fields = ("H", "f", "G", "h", "A", "b")

    def __init__(self, nz, nl, nv, max_batch=1, device=0):
        super().__init__()
        if nz <= 0 or nv <= 0 or nl < 0:  # fbstab_dense.cc:19-23
            raise RuntimeError("In FBstabDense::FBstabDense: nz and nv must be "
                               "positive, nl nonnegative")
        self.nz, self.nl, self.nv, self.max_batch = nz, nl, nv, max_batch
        self.device = device
        self.field_sizes = {"H": nz * nz, "f": nz, "G": nl * nz, "h": nl,
                            "A": nv * nz, "b": nv}
        capi.check(capi.lib().fbstab_dense_batch_create(
            nz, nl, nv, max_batch, device, C.byref(self._h)))

    def solve(self, H, f, G, h, A, b, x0=None):
        """One instance from 2-D numpy matrices (any memory order).
        Returns (out record, (z, l, v, y))."""
        cm = lambda M, r, c: np.ascontiguousarray(  # column-major flatten
            np.asarray(M, dtype=np.float64).reshape(r, c).T).reshape(-1)
        vec = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1)
        nz, nl, nv = self.nz, self.nl, self.nv
        f_, h_, b_ = vec(f), vec(h), vec(b)
        if f_.size != nz or h_.size != nl or b_.size != nv:
            raise RuntimeError("In FBstabDense::Solve: mismatch between *this "
                               "and data dimensions.")
        data = {"H": cm(H, nz, nz), "f": f_, "G": cm(G, nl, nz) if nl else np.zeros(0),
                "h": h_, "A": cm(A, nv, nz), "b": b_}
        if x0 is None:
            z, l, v = np.zeros(nz), np.zeros(nl), np.zeros(nv)
        else:
            z, l, v = [vec(t).copy() for t in x0]
            if z.size != nz or l.size != nl or v.size != nv:
                raise RuntimeError("In FBstabDense::Solve: mismatch between "
                                   "*this and initial guess dimensions.")
        out, y = self.solve_batch(data, z, l, v)
        return out[0], (z, l, v, y)


class FBstabMpc(_Base):
    """Linear-quadratic OCP in the reference's stage form (fbstab_mpc.h:17-55)."""
    _prefix = "mpc"
    _fields = ("Q", "R", "S", "q", "r", "A", "B", "c", "E", "L", "d", "x0")

    def __init__(self, N, nx, nu, nc, max_batch=1, device=0):
        super().__init__()
        if N < 1 or nx < 1 or nu < 1 or nc < 1:  # fbstab_mpc.cc:62-65
            raise RuntimeError(
                "In FBstabMpc::FBstabMpc: problem sizes must be positive.")
        self.N, self.nx, self.nu, self.nc = N, nx, nu, nc
        self.max_batch = max_batch
        self.device = device
        K = N + 1
        self.nz, self.nl, self.nv = K * (nx + nu), K * nx, K * nc
        self.field_sizes = {"Q": K * nx * nx, "R": K * nu * nu, "S": K * nu * nx,
                            "q": K * nx, "r": K * nu, "A": N * nx * nx,
                            "B": N * nx * nu, "c": N * nx, "E": K * nc * nx,
                            "L": K * nc * nu, "d": K * nc, "x0": nx}
        capi.check(capi.lib().fbstab_mpc_batch_create(
            N, nx, nu, nc, max_batch, device, C.byref(self._h)))

    def _solve_one_copy(self, fn_name, sizes, data, z, l, v, y, out, stream):
        batch = (z.size if _is_host(z) else z.numel()) // self.nz
        if batch > self.max_batch:
            raise RuntimeError("batch exceeds max_batch")
        _flat(z, batch * self.nz, "z", self.device)
        _flat(l, batch * self.nl, "l", self.device)
        _flat(v, batch * self.nv, "v", self.device)
        if y is None:
            y = np.zeros(batch * self.nv) if _is_host(z) else z.new_zeros(batch * self.nv)
        if out is None:
            if _is_host(z):
                out = np.zeros(batch, dtype=OUT_DTYPE)
            else:
                import torch
                out = torch.zeros(batch * OUT_DTYPE.itemsize, dtype=torch.uint8, device=z.device)
        ptrs = [capi.ptr(_flat(data[k], (batch if k == "x0" else 1) * sizes[k], k, self.device))
                for k in self._fields]
        fn = getattr(capi.lib(), fn_name)
        fn.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 18
        capi.check(fn(self._h, batch, *ptrs, capi.ptr(z), capi.ptr(l), capi.ptr(v), capi.ptr(y),
                      capi.ptr(out), stream))
        return out, y

    def solve_batch_shared(self, data, z, l, v, y=None, out=None, stream=None):
        """ONE copy of the 11 stage-data sequences for the whole batch; data["x0"]
        holds the `batch` initial states (fbstab_mpc_batch_solve_shared)."""
        return self._solve_one_copy("fbstab_mpc_batch_solve_shared", self.field_sizes, data,
                                    z, l, v, y, out, stream)

    def solve_batch_lti(self, data, z, l, v, y=None, out=None, stream=None):
        """ONE STAGE of each sequence (time-invariant problem), replicated over the
        horizon like OcpGenerator::CopyOverHorizon (fbstab_mpc_batch_solve_lti)."""
        nx, nu, nc = self.nx, self.nu, self.nc
        one = {"Q": nx * nx, "R": nu * nu, "S": nu * nx, "q": nx, "r": nu, "A": nx * nx,
               "B": nx * nu, "c": nx, "E": nc * nx, "L": nc * nu, "d": nc, "x0": nx}
        return self._solve_one_copy("fbstab_mpc_batch_solve_lti", one, data, z, l, v, y, out,
                                    stream)

    def solve(self, data, x0=None):
        """One instance; `data` as produced by problems.ocp_batch(count=1)."""
        if x0 is None:
            z, l, v = np.zeros(self.nz), np.zeros(self.nl), np.zeros(self.nv)
        else:
            z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy()
                       for t in x0]
        out, y = self.solve_batch(data, z, l, v)
        return out[0], (z, l, v, y)


class FBstabSparse(_Base):
    """min 1/2 z'Hz + f'z  s.t. Gz = h, Az <= b with SPARSE H, G, A; every instance of a
    batch shares one sparsity pattern (the reference plans this solver: ROADMAP.md:10,
    over the LDL' wrapper tools/qdldl/qdldl_wrapper.h:19-84).

    pattern = (Hp, Hi, Gp, Gi, Ap, Ai): H (nz x nz) by its upper triangle, G (nl x nz) and
    A (nv x nz) in compressed-column form, row indices increasing within a column.
    The constructor runs the symbolic analysis (elimination order, pattern of L)."""
    _prefix = "sparse"
    _fields = ("Hx", "f", "Gx", "h", "Ax", "b")

    def __init__(self, nz, nl, nv, pattern, max_batch=1, device=0, perm=None):
        super().__init__()
        if nz <= 0 or nv <= 0 or nl < 0:  # as FBstabDense, fbstab_dense.cc:19-23
            raise RuntimeError("In FBstabSparse::FBstabSparse: nz and nv must be "
                               "positive, nl nonnegative")
        ia = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
        self.pattern = tuple(ia(a) for a in pattern)
        Hp, Hi, Gp, Gi, Ap, Ai = self.pattern
        if Hp.size != nz + 1 or Ap.size != nz + 1 or (nl > 0 and Gp.size != nz + 1):
            raise RuntimeError("In FBstabSparse::FBstabSparse: column pointer arrays must "
                               "have nz + 1 entries")
        if Hi.size != Hp[-1] or Ai.size != Ap[-1] or (nl > 0 and Gi.size != Gp[-1]):
            raise RuntimeError("In FBstabSparse::FBstabSparse: index arrays do not match "
                               "the column pointers")
        self.nz, self.nl, self.nv, self.max_batch = nz, nl, nv, max_batch
        self.device = device
        self.field_sizes = {"Hx": int(Hp[-1]), "f": nz, "Gx": int(Gp[-1]) if nl > 0 else 0,
                            "h": nl, "Ax": int(Ap[-1]), "b": nv}
        pm = ia(perm) if perm is not None else None
        if pm is not None and pm.size != nz + nl + nv:
            raise RuntimeError("In FBstabSparse::FBstabSparse: perm must have nz + nl + nv entries")
        capi.check(capi.lib().fbstab_sparse_batch_create(
            nz, nl, nv, capi.ptr(Hp), capi.ptr(Hi), capi.ptr(Gp) if nl > 0 else None,
            capi.ptr(Gi) if nl > 0 and Gi.size else None, capi.ptr(Ap), capi.ptr(Ai),
            capi.ptr(pm), max_batch, device, C.byref(self._h)))

    def solve_batch_devices(self, data, z, l, v, devices, perm=None):
        """The batch cut into contiguous ranges over several GPUs of this box, one host
        thread and one engine handle per device (fbstab_sparse_multi_gpu_solve): HOST
        arrays only; returns (out, y) like solve_batch, identical byte for byte."""
        L = capi.lib()
        L.fbstab_sparse_multi_gpu_create.argtypes = (
            [C.c_int, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] * 7 +
            [C.c_long, C.POINTER(C.c_void_p)])
        L.fbstab_sparse_multi_gpu_solve.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 11
        L.fbstab_sparse_multi_gpu_set_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        L.fbstab_sparse_multi_gpu_destroy.argtypes = [C.c_void_p]
        if not all(_is_host(a) for a in (z, l, v)):
            raise RuntimeError("solve_batch_devices takes host arrays")
        batch = z.size // self.nz
        Hp, Hi, Gp, Gi, Ap, Ai = self.pattern
        pm = np.ascontiguousarray(np.asarray(perm, dtype=np.int32)) if perm is not None else None
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        capi.check(L.fbstab_sparse_multi_gpu_create(
            len(devices), devs, self.nz, self.nl, self.nv, capi.ptr(Hp), capi.ptr(Hi),
            capi.ptr(Gp) if self.nl > 0 else None,
            capi.ptr(Gi) if self.nl > 0 and Gi.size else None, capi.ptr(Ap), capi.ptr(Ai),
            capi.ptr(pm), max(batch, 1), C.byref(h)))
        try:
            if getattr(self, "opts", None) is not None:
                capi.check(L.fbstab_sparse_multi_gpu_set_options(h, C.byref(self.opts)))
            y = np.zeros(batch * self.nv)
            out = np.zeros(batch, dtype=OUT_DTYPE)
            capi.check(L.fbstab_sparse_multi_gpu_solve(
                h, batch, *[capi.ptr(_flat(data[k], batch * self.field_sizes[k], k, None))
                            for k in self._fields],
                capi.ptr(_flat(z, batch * self.nz, "z", None)),
                capi.ptr(_flat(l, batch * self.nl, "l", None)),
                capi.ptr(_flat(v, batch * self.nv, "v", None)), capi.ptr(y), capi.ptr(out)))
        finally:
            L.fbstab_sparse_multi_gpu_destroy(h)
        return out, y

    @staticmethod
    def analyze(nz, nl, nv, pattern, perm=None):
        """(n, nnz(K), nnz(L), perm) of the symbolic analysis the constructor would run for
        this pattern -- host only, no device needed (fbstab_sparse_analyze)."""
        ia = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
        Hp, Hi, Gp, Gi, Ap, Ai = (ia(a) for a in pattern)
        pm = ia(perm) if perm is not None else None
        n, k, l = C.c_int(), C.c_int(), C.c_int()
        out = np.zeros(nz + nl + nv, dtype=np.int32)
        capi.check(capi.lib().fbstab_sparse_analyze(
            nz, nl, nv, capi.ptr(Hp), capi.ptr(Hi), capi.ptr(Gp) if nl > 0 else None,
            capi.ptr(Gi) if nl > 0 and Gi.size else None, capi.ptr(Ap), capi.ptr(Ai),
            capi.ptr(pm), C.byref(n), C.byref(k), C.byref(l), capi.ptr(out)))
        return n.value, k.value, l.value, out

    def analysis(self):
        """(n, nnz(K), nnz(L), perm) of the symbolic analysis; perm[new] = old over [z; l; w]."""
        n, k, l = C.c_int(), C.c_int(), C.c_int()
        capi.check(capi.lib().fbstab_sparse_batch_analysis(
            self._h, C.byref(n), C.byref(k), C.byref(l), None))
        perm = np.zeros(n.value, dtype=np.int32)
        capi.check(capi.lib().fbstab_sparse_batch_analysis(
            self._h, None, None, None, capi.ptr(perm)))
        return n.value, k.value, l.value, perm

    def factor_pattern(self):
        """(Lp, Li): the strictly lower triangle of L, compressed columns."""
        n, _, nnzL, _ = self.analysis()
        Lp, Li = np.zeros(n + 1, dtype=np.int32), np.zeros(nnzL, dtype=np.int32)
        capi.check(capi.lib().fbstab_sparse_batch_factor_pattern(
            self._h, capi.ptr(Lp), capi.ptr(Li) if nnzL else None))
        return Lp, Li

    def component(self, *a, **k):
        raise RuntimeError("FBstabSparse has no component-stage entry")
