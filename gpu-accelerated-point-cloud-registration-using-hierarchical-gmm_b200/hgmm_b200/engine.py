"""Engine: thin object wrapper over the libhgmm C ABI (one context = one device + one stream)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and hasattr(x, "data_ptr") and bool(x.is_cuda)


class Engine:
    """Owns an `hgmm_ctx`.  `stream` may be an int cudaStream_t handle (e.g. torch.cuda.Stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self._lib = L.load()
        self._ctx = C.c_void_p()
        handle = C.c_void_p(int(stream)) if stream else None
        rc = self._lib.hgmm_create(C.byref(self._ctx), int(device), handle)
        if rc != L.HGMM_OK or not self._ctx:
            raise L.HgmmError("hgmm_create failed (status %d): no usable CUDA device %d -- libhgmm has no CPU path" % (rc, device))
        self.device = int(device)
        self._stream = int(stream) if stream else None
        self.n_points = 0
        self.model_token = None       # who installed the tree / target last (GMMTree instances share engines safely through it)
        self.target_token = None

    # -- plumbing ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.hgmm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != L.HGMM_OK:
            msg = self._lib.hgmm_last_error(self._ctx)
            raise L.HgmmError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))

    @property
    def launch_count(self):
        return int(self._lib.hgmm_launch_count(self._ctx))

    @property
    def total_points(self):
        return int(self._lib.hgmm_total_points(self._ctx))

    def last_timing_ms(self):
        out = np.zeros(3)
        self._check(self._lib.hgmm_last_timing(self._ctx, L.ptr(out)), "hgmm_last_timing")
        return out

    def set_profiling(self, on):
        self._check(self._lib.hgmm_set_profiling(self._ctx, int(bool(on))), "hgmm_set_profiling")

    def measure_fp32_peak(self):
        """-> TFLOP/s of (immediate-operand FFMA, 3-register FFMA, packed FFMA2)"""
        out = np.zeros(3)
        self._check(self._lib.hgmm_measure_fp32_peak(self._ctx, L.ptr(out)), "hgmm_measure_fp32_peak")
        return float(out[0]), float(out[1]), float(out[2])

    def sorted_points(self):
        """-> [n,3] float32: the cloud in the Morton-cell order the J > 512 sweep reads it (csrc/cloud_sort.cu)"""
        out = np.empty((self.n_points, 3), np.float32)
        self._check(self._lib.hgmm_sorted_points(self._ctx, L.ptr(out)), "hgmm_sorted_points")
        return out

    @staticmethod
    def _cloud_arg(points):
        """-> (void*, n, mem_kind, keepalive). Accepts numpy [N,3] (host) or a CUDA torch tensor [N,3] fp32 contiguous."""
        if _is_torch_cuda(points):
            t = points
            if t.dim() != 2 or t.shape[1] != 3 or str(t.dtype) != "torch.float32" or not t.is_contiguous():
                raise ValueError("device clouds must be contiguous float32 [N,3]")
            # libhgmm reads the tensor on ITS stream: the stream that produced it must be done first (include/hgmm.h)
            import torch
            torch.cuda.current_stream(t.device).synchronize()
            return C.c_void_p(t.data_ptr()), int(t.shape[0]), L.MEM_DEVICE, t
        if hasattr(points, "is_pinned") and hasattr(points, "numpy"):      # CPU torch tensor (possibly pinned)
            t = points
            if t.dim() != 2 or t.shape[1] != 3 or str(t.dtype) != "torch.float32" or not t.is_contiguous():
                raise ValueError("host tensors must be contiguous float32 [N,3]")
            return C.c_void_p(t.data_ptr()), int(t.shape[0]), L.MEM_HOST, t
        a = np.asarray(points.points if hasattr(points, "points") else points)
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError("point cloud must be [N,3]")
        a = L.f32c(a)
        return L.ptr(a), int(a.shape[0]), L.MEM_HOST, a

    # -- data --------------------------------------------------------------------------------
    def set_points(self, points, total=None):
        """`total` (multi-GPU): the number of points over all ranks when the caller knows it -- skips the all-reduce."""
        p, n, kind, keep = self._cloud_arg(points)
        if total is not None:
            self.declare_total_points(total)
        self._check(self._lib.hgmm_set_points(self._ctx, p, n, kind), "hgmm_set_points")
        self._src_keep = keep         # a pinned host tensor is read asynchronously: hold it until the next cloud replaces it
        self.n_points = n
        self.model_token = None
        return self

    def declare_total_points(self, total):
        self._check(self._lib.hgmm_declare_total_points(self._ctx, int(total)), "hgmm_declare_total_points")

    # -- flat mixture ------------------------------------------------------------------------
    def fit_flat(self, means, covs, weights, cov_type="full", flavor=None, max_iter=10, tol=0.0, sigma_bug=False,
                 tile_points=0, want_outputs=True, variant=0):
        ct = L.COV_TYPES[cov_type]
        if flavor is None:
            flavor = L.FLAVOR_CPP if ct == L.COV_FULL else L.FLAVOR_PY
        means = L.f32c(means)
        J = means.shape[0]
        ce = {L.COV_FULL: (J, 3, 3), L.COV_DIAG: (J, 3), L.COV_SPHERICAL: (J,)}[ct]
        covs = L.f32c(covs, ce)
        weights = L.f32c(weights, (J,))
        cfg = L.FlatConfig(J, ct, flavor, int(max_iter), float(tol), int(bool(sigma_bug)), int(tile_points), int(variant))
        o_means = np.empty((J, 3), np.float32) if want_outputs else None
        o_covs = np.empty(ce, np.float32) if want_outputs else None
        o_w = np.empty(J, np.float32) if want_outputs else None
        o_inv = np.empty((J, 3) if ct == L.COV_DIAG else (J,), np.float32) if (want_outputs and flavor != L.FLAVOR_CPP) else None
        o_ll = np.zeros(max(int(max_iter), 1), np.float64)
        o_it = np.zeros(1, np.int32)
        rc = self._lib.hgmm_fit_flat(self._ctx, C.byref(cfg), L.ptr(means), L.ptr(covs), L.ptr(weights), L.ptr(o_means),
                                     L.ptr(o_covs), L.ptr(o_w), L.ptr(o_inv), L.ptr(o_ll), L.ptr(o_it))
        self._check(rc, "hgmm_fit_flat")
        it = int(o_it[0])
        return {"means": o_means, "covs": o_covs, "weights": o_w, "inv_cov": o_inv, "ll": o_ll[:it], "iters": it}

    def predict_flat(self, points=None):
        if points is None:
            n = self.n_points
            labels = np.empty(n, np.int32)
            rc = self._lib.hgmm_predict_flat(self._ctx, None, 0, L.MEM_HOST, L.ptr(labels))
        else:
            p, n, kind, keep = self._cloud_arg(points)
            labels = np.empty(n, np.int32)
            rc = self._lib.hgmm_predict_flat(self._ctx, p, n, kind, L.ptr(labels))
        self._check(rc, "hgmm_predict_flat")
        return labels

    # -- hierarchical mixture ----------------------------------------------------------------
    @staticmethod
    def tree_total_nodes(max_level):
        return int(L.load().hgmm_tree_total_nodes(int(max_level)))

    def fit_tree(self, init_means, max_level, ls=20.0, ld=1.0e-4, sig2=0.004, ll_mode="level", max_iters_per_level=10000,
                 chunk_points=0, want_current=True, want_outputs=True, variant=0, prune_lambda_c=0.0, prune_min_points=0.0):
        """prune_lambda_c / prune_min_points > 0: adaptive (ragged) build, see include/hgmm.h (ll_mode must be 'estep')"""
        nt = self.tree_total_nodes(max_level)
        init_means = L.f32c(init_means, (nt, 3))
        cfg = L.TreeConfig(int(max_level), L.LL_LEVEL if ll_mode == "level" else L.LL_ESTEP, float(ls), float(ld), float(sig2),
                           int(max_iters_per_level), int(chunk_points), int(variant), float(prune_lambda_c), float(prune_min_points))
        pi = np.empty(nt, np.float32) if want_outputs else None
        mu = np.empty((nt, 3), np.float32) if want_outputs else None
        cov = np.empty((nt, 3, 3), np.float32) if want_outputs else None
        cur = np.empty(self.n_points, np.int64) if want_current else None
        iters = np.zeros(max_level, np.int32)
        q = np.zeros(max_level, np.float64)
        rc = self._lib.hgmm_fit_tree(self._ctx, C.byref(cfg), L.ptr(init_means), L.ptr(pi), L.ptr(mu), L.ptr(cov), L.ptr(cur),
                                     L.ptr(iters), L.ptr(q))
        self._check(rc, "hgmm_fit_tree")
        self.model_token = None
        return {"pi": pi, "mu": mu, "cov": cov, "current": cur, "iters": iters, "q": q}

    def tree_set_model(self, max_level, pi, mu, cov):
        nt = self.tree_total_nodes(max_level)
        pi, mu, cov = L.f32c(pi, (nt,)), L.f32c(mu, (nt, 3)), L.f32c(cov, (nt, 3, 3))
        self._check(self._lib.hgmm_tree_set_model(self._ctx, int(max_level), L.ptr(pi), L.ptr(mu), L.ptr(cov)), "hgmm_tree_set_model")
        self._tree_level = int(max_level)
        self.model_token = None
        return self

    # -- registration ------------------------------------------------------------------------
    def reg_set_target(self, target):
        p, n, kind, keep = self._cloud_arg(target)
        self._check(self._lib.hgmm_reg_set_target(self._ctx, p, n, kind), "hgmm_reg_set_target")
        self._tgt_keep = keep
        self.target_token = None
        return self

    def reg_estep(self, rot, t, lambda_c, nt, want_m2=True):
        rot = np.ascontiguousarray(rot, np.float64).reshape(3, 3)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        m0 = np.zeros(nt)
        m1 = np.zeros((nt, 3))
        m2 = np.zeros((nt, 3, 3)) if want_m2 else None
        rc = self._lib.hgmm_reg_estep(self._ctx, L.ptr(rot), L.ptr(t), float(lambda_c), L.ptr(m0), L.ptr(m1), L.ptr(m2))
        self._check(rc, "hgmm_reg_estep")
        return m0, m1, m2

    def reg_mstep(self, rot, t, solver="twist_lstsq"):
        rot = np.array(rot, np.float64).reshape(3, 3).copy()
        t = np.array(t, np.float64).reshape(3).copy()
        q = np.zeros(1)
        rc = self._lib.hgmm_reg_mstep(self._ctx, L.SOLVERS[solver], L.ptr(rot), L.ptr(t), L.ptr(q))
        self._check(rc, "hgmm_reg_mstep")
        return rot, t, float(q[0])

    def register_tree(self, rot=None, t=None, solver="twist_lstsq", maxiter=20, tol=1.0e-4, lambda_c=0.01):
        rot = np.identity(3) if rot is None else np.array(rot, np.float64).reshape(3, 3).copy()
        t = np.zeros(3) if t is None else np.array(t, np.float64).reshape(3).copy()
        cfg = L.RegConfig(L.SOLVERS[solver], int(maxiter), float(tol), float(lambda_c))
        q = np.zeros(1)
        it = np.zeros(1, np.int32)
        hist = np.zeros(int(maxiter))
        rc = self._lib.hgmm_register_tree(self._ctx, C.byref(cfg), L.ptr(rot), L.ptr(t), L.ptr(q), L.ptr(it), L.ptr(hist))
        self._check(rc, "hgmm_register_tree")
        return rot, t, float(q[0]), int(it[0]), hist[:int(it[0])]

    def register_flat(self, rot=None, t=None, solver="procrustes_svd", maxiter=20, tol=1.0e-4):
        """registration of the target against the flat mixture of the last fit_flat -> (rot, t, q, iterations, q history);
        (rot, t) is the FORWARD transform (target -> model frame), like register_tree"""
        rot = np.identity(3) if rot is None else np.array(rot, np.float64).reshape(3, 3).copy()
        t = np.zeros(3) if t is None else np.array(t, np.float64).reshape(3).copy()
        cfg = L.RegConfig(L.SOLVERS[solver], int(maxiter), float(tol), 0.0)
        q = np.zeros(1)
        it = np.zeros(1, np.int32)
        hist = np.zeros(int(maxiter))
        rc = self._lib.hgmm_register_flat(self._ctx, C.byref(cfg), L.ptr(rot), L.ptr(t), L.ptr(q), L.ptr(it), L.ptr(hist))
        self._check(rc, "hgmm_register_flat")
        return rot, t, float(q[0]), int(it[0]), hist[:int(it[0])]

    # -- L2 registration of two flat mixtures ---------------------------------------------------
    def l2_set_mixtures(self, mu_source, phi_source, mu_target, phi_target):
        ms = np.ascontiguousarray(mu_source, np.float64).reshape(-1, 3)
        ps = np.ascontiguousarray(phi_source, np.float64).reshape(-1)
        mt = np.ascontiguousarray(mu_target, np.float64).reshape(-1, 3)
        pt = np.ascontiguousarray(phi_target, np.float64).reshape(-1)
        if len(ms) != len(ps) or len(mt) != len(pt):
            raise ValueError("means / weights length mismatch")
        rc = self._lib.hgmm_l2_set_mixtures(self._ctx, L.ptr(ms), L.ptr(ps), len(ps), L.ptr(mt), L.ptr(pt), len(pt))
        self._check(rc, "hgmm_l2_set_mixtures")
        return self

    def l2_cost_grad(self, theta, sigma):
        th = np.ascontiguousarray(theta, np.float64).reshape(7)
        f = np.zeros(1)
        g = np.zeros(7)
        self._check(self._lib.hgmm_l2_cost_grad(self._ctx, L.ptr(th), float(sigma), L.ptr(f), L.ptr(g)), "hgmm_l2_cost_grad")
        return float(f[0]), g

    def l2_optimize(self, theta, sigma, max_iter=10, gtol=1.0e-5):
        """-> (theta, f, iterations, evaluations, status)"""
        th = np.array(theta, np.float64).reshape(7).copy()
        f = np.zeros(1)
        it, nf, st = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32)
        rc = self._lib.hgmm_l2_optimize(self._ctx, L.ptr(th), float(sigma), int(max_iter), float(gtol), L.ptr(f), L.ptr(it),
                                        L.ptr(nf), L.ptr(st))
        self._check(rc, "hgmm_l2_optimize")
        return th, float(f[0]), int(it[0]), int(nf[0]), int(st[0])

    def fill_vbo(self, vbo_pos_devptr, vbo_col_devptr, scene_scale=0.1):
        rc = self._lib.hgmm_fill_vbo(self._ctx, C.c_void_p(vbo_pos_devptr) if vbo_pos_devptr else None,
                                     C.c_void_p(vbo_col_devptr) if vbo_col_devptr else None, float(scene_scale), None, None)
        self._check(rc, "hgmm_fill_vbo")

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_init(self, rank, nranks, unique_id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id_bytes))
        self._check(self._lib.hgmm_comm_init(self._ctx, int(rank), int(nranks), C.cast(buf, C.c_void_p)), "hgmm_comm_init")

    def p2p_export(self):
        """this rank's exchange-window handle (64 bytes, cudaIpcMemHandle_t)"""
        buf = (C.c_char * 64)()
        self._check(self._lib.hgmm_p2p_export(self._ctx, C.cast(buf, C.c_void_p)), "hgmm_p2p_export")
        return bytes(buf)

    def p2p_attach(self, handles):
        """handles: one 64-byte handle per rank, in rank order"""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != 64 * len(handles):
            raise ValueError("every handle must be 64 bytes")
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.hgmm_p2p_attach(self._ctx, C.cast(buf, C.c_void_p), len(handles)), "hgmm_p2p_attach")

    def p2p_detach(self):
        self._check(self._lib.hgmm_p2p_detach(self._ctx), "hgmm_p2p_detach")

    @property
    def p2p_enabled(self):
        return bool(self._lib.hgmm_p2p_enabled(self._ctx))

    def comm_destroy(self):
        self._check(self._lib.hgmm_comm_destroy(self._ctx), "hgmm_comm_destroy")

    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = L.load().hgmm_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != L.HGMM_OK:
            raise L.HgmmError("hgmm_comm_unique_id failed (status %d): NCCL not loadable" % rc)
        return bytes(buf)
