# cython: language_level=3, boundscheck=False, wraparound=False
# distutils: language = c++
"""Python surface of pydensecrf, bound to the C-ABI of libdcrf_b200.so (include/dcrf_b200.h).

Mirrors the classes, signatures, argument meaning and error behaviour of the reference's Cython
layer (pydensecrf/densecrf.pyx and pydensecrf/eigen.pyx in /root/reference), but every method is a
thin call into the C-ABI; no Eigen, no C++ classes, no CPU compute path.  `pydensecrf_b200.densecrf`
and `pydensecrf_b200.eigen` re-export the names under the reference's module layout.
"""
from numbers import Number
import enum

from cpython.buffer cimport PyBUF_WRITABLE
from libc.stdint cimport int16_t, int32_t, int64_t, uint8_t, uint64_t
from libc.stdlib cimport malloc, free


cdef extern from "dcrf_b200.h" nogil:
    ctypedef struct dcrf_t:
        pass
    ctypedef struct dcrf_mat_t:
        pass
    ctypedef struct dcrf_prof_entry:
        char name[32]
        int launches
        float ms_total
    const char* dcrf_last_error()
    int dcrf_abi_version()
    int dcrf_device_count()
    int dcrf_create(int nvar, int nlabels, int device, dcrf_t** out)
    int dcrf_create_2d(int width, int height, int nlabels, int device, dcrf_t** out)
    void dcrf_destroy(dcrf_t* m)
    int dcrf_set_unary_const(dcrf_t* m, const float* u)
    int dcrf_set_unary_logistic(dcrf_t* m, const float* Lmat, int F, const float* f)
    int dcrf_set_unary_from_softmax(dcrf_t* m, const float* sm, double scale, double clip)
    int dcrf_set_unary_from_labels(dcrf_t* m, const int32_t* labels, double gt_prob, int zero_unsure)
    int dcrf_set_unary_device(dcrf_t* m, const float* dev_u)
    int dcrf_mat_device_ptr(dcrf_mat_t* a, float** ptr, int* pitch, int* device)
    int dcrf_add_pairwise_nd(dcrf_t* m, int nsp, const int* dims, const float* sdims, int nch, const float* schan,
                             const void* img, int img_is_float, int ckind, const float* cdata, int ktype, int ntype)
    int dcrf_add_pairwise_energy(dcrf_t* m, const float* feat, int d, int ckind, const float* cdata, int ktype, int ntype)
    int dcrf_add_pairwise_gaussian(dcrf_t* m, float sx, float sy, int ckind, const float* cdata, int ktype, int ntype)
    int dcrf_add_pairwise_bilateral(dcrf_t* m, float sx, float sy, float sr, float sg, float sb, const uint8_t* rgb,
                                    int ckind, const float* cdata, int ktype, int ntype)
    int dcrf_inference(dcrf_t* m, int n, dcrf_mat_t** Q)
    int dcrf_inference_into(dcrf_t* m, int n, dcrf_mat_t* Q)
    int dcrf_start_inference(dcrf_t* m, dcrf_mat_t** Q)
    int dcrf_step_inference(dcrf_t* m, dcrf_mat_t* Q, dcrf_mat_t* t1, dcrf_mat_t* t2)
    int dcrf_kl_divergence(dcrf_t* m, dcrf_mat_t* Q, double* out)
    int dcrf_current_map(dcrf_t* m, dcrf_mat_t* Q, int16_t* labels)
    int dcrf_map(dcrf_t* m, int n, int16_t* labels)
    int dcrf_mat_create_empty(dcrf_mat_t** out)
    int dcrf_mat_from_host(const float* rowmajor, int rows, int cols, dcrf_mat_t** out)
    int dcrf_vec_from_host(const float* data, int n, dcrf_mat_t** out)
    void dcrf_mat_destroy(dcrf_mat_t* a)
    int dcrf_mat_rows(const dcrf_mat_t* a)
    int dcrf_mat_cols(const dcrf_mat_t* a)
    int dcrf_mat_on_device(const dcrf_mat_t* a)
    int dcrf_mat_host_ptr(dcrf_mat_t* a, float** ptr)
    void dcrf_mat_release_host(dcrf_mat_t* a)
    int dcrf_synchronize(dcrf_t* m)
    int dcrf_timer_start(dcrf_t* m)
    int dcrf_timer_stop(dcrf_t* m, float* ms)
    int dcrf_profile_enable(dcrf_t* m, int on)
    int dcrf_profile_read(dcrf_t* m, dcrf_prof_entry* out, int cap, int* n, int reset)
    int64_t dcrf_launch_count(const dcrf_t* m)
    int dcrf_flush_l2(dcrf_t* m)
    int dcrf_mat_argmax(dcrf_mat_t* a, void* labels, int label_bytes)
    int dcrf_debug_selftest(dcrf_t* m, int which, uint64_t n, uint64_t seed, uint64_t* mismatches)
    int dcrf_batch_map_2d(int W, int H, int L, int device, int B, const float* const* unary, const uint8_t* const* rgb,
                          float g_sx, float g_sy, float g_w, float b_sx, float b_sy, float b_sr, float b_sg, float b_sb,
                          float b_w, int n_iterations, int n_workers, int16_t* const* labels, float* const* q_out,
                          const int* devices, int n_devices)
    int dcrf_num_pairwise(const dcrf_t* m)
    int dcrf_lattice_info(const dcrf_t* m, int k, int* d, int* N, int* V)
    int dcrf_lattice_dump(dcrf_t* m, int k, int32_t* offset, float* bary, int16_t* keys, int32_t* nbr, float* norm)
    int dcrf_lattice_filter(dcrf_t* m, int k, const float* inp, int L, int reverse, float* out)
    int dcrf_pairwise_apply(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out)
    int dcrf_pairwise_apply_transpose(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out)
    int dcrf_lattice_gradient(dcrf_t* m, int k, const float* a, const float* b, int L, float* df)
    ctypedef struct dcrf_objective:
        int kind
        const int16_t* gt
        float robust
        const float* class_weight
        int n_class_weight
    int dcrf_gradient(dcrf_t* m, int n_iterations, const dcrf_objective* objective, double* value, float* unary_grad,
                      float* lbl_cmp_grad, float* kernel_grad)
    int dcrf_unary_energy(dcrf_t* m, const int16_t* labels, float* out)
    int dcrf_pairwise_energy(dcrf_t* m, const int16_t* labels, int term, float* out)
    int dcrf_num_parameters(dcrf_t* m, int group, int* n)
    int dcrf_get_parameters(dcrf_t* m, int group, float* out)
    int dcrf_set_parameters(dcrf_t* m, int group, const float* v, int n)
    int dcrf_set_iteration_mode(dcrf_t* m, int mode)
    int dcrf_set_softmax_mode(dcrf_t* m, int fast)
    int dcrf_set_kernel_variant(dcrf_t* m, int mask)
    int dcrf_set_kl_mode(dcrf_t* m, int mode)
    void dcrf_empty_cache()
    int dcrf_set_matrix_mode(dcrf_t* m, int mode)
    int dcrf_host_alloc(size_t nbytes, void** out)
    int dcrf_host_alloc_result(size_t nbytes, void** out)
    void dcrf_host_free(void* p, size_t nbytes)


class KernelType(enum.IntEnum):
    # pydensecrf/densecrf/include/pairwise.h:38-42
    CONST_KERNEL = 0
    DIAG_KERNEL = 1
    FULL_KERNEL = 2


class NormalizationType(enum.IntEnum):
    # pydensecrf/densecrf/include/pairwise.h:32-37
    NO_NORMALIZATION = 0
    NORMALIZE_BEFORE = 1
    NORMALIZE_AFTER = 2
    NORMALIZE_SYMMETRIC = 3


# the README and examples/inference.py use the module-level spellings (dcrf.DIAG_KERNEL, ...)
CONST_KERNEL = KernelType.CONST_KERNEL
DIAG_KERNEL = KernelType.DIAG_KERNEL
FULL_KERNEL = KernelType.FULL_KERNEL
NO_NORMALIZATION = NormalizationType.NO_NORMALIZATION
NORMALIZE_BEFORE = NormalizationType.NORMALIZE_BEFORE
NORMALIZE_AFTER = NormalizationType.NORMALIZE_AFTER
NORMALIZE_SYMMETRIC = NormalizationType.NORMALIZE_SYMMETRIC


cdef int _check(int status) except -1:
    if status == 0:
        return 0
    msg = dcrf_last_error().decode("utf-8", "replace")
    if status == 1:
        raise ValueError(msg)
    if status == 3:
        raise MemoryError(msg)
    raise RuntimeError(msg)


cdef class _PinnedBlock:
    """Owner of one page-locked host block; exposes it as a writable byte buffer."""
    cdef void* p
    cdef size_t n
    cdef Py_ssize_t shape[1]
    cdef Py_ssize_t strides[1]

    def __cinit__(self, size_t nbytes, bint result=False):
        self.p = NULL
        self.n = nbytes
        if result:   # may leave p NULL: too many results of this size are alive, the caller takes ordinary memory
            _check(dcrf_host_alloc_result(nbytes, &self.p))
        else:
            _check(dcrf_host_alloc(nbytes, &self.p))
        self.shape[0] = nbytes
        self.strides[0] = 1

    def __dealloc__(self):
        if self.p != NULL:
            dcrf_host_free(self.p, self.n)
            self.p = NULL

    def __getbuffer__(self, Py_buffer* buf, int flags):
        buf.buf = <char*>self.p
        buf.format = 'B'
        buf.internal = NULL
        buf.itemsize = 1
        buf.len = self.n
        buf.ndim = 1
        buf.obj = self
        buf.readonly = 0
        buf.shape = self.shape
        buf.strides = self.strides
        buf.suboffsets = NULL

    def __releasebuffer__(self, Py_buffer* buf):
        pass


def pinned_empty(shape, dtype="float32"):
    """numpy array backed by page-locked host memory (needs a CUDA device).  Unaries / images placed in
    such arrays are uploaded by DMA at full PCIe rate instead of through the driver's pageable staging."""
    import numpy as np
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    blk = _PinnedBlock(max(n, 1))
    return np.frombuffer(blk, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def _result_empty(shape, dtype):
    """Array for a result the caller may keep (label maps): page-locked from the pool while few of its size are alive,
    ordinary numpy memory otherwise -- pinning a fresh block per call costs more than the copy it speeds up."""
    import numpy as np
    dt = np.dtype(dtype)
    cdef Py_ssize_t cnt = int(np.prod(shape))
    cdef _PinnedBlock blk = _PinnedBlock(max(cnt * dt.itemsize, 1), True)
    if blk.p == NULL:
        return np.empty(shape, dt)
    return np.frombuffer(blk, dtype=dt, count=cnt).reshape(shape)


def abi_version():
    return dcrf_abi_version()


def device_count():
    """Number of CUDA devices the engine can see (0 without a GPU/driver)."""
    return dcrf_device_count()


# ------------------------------------------------------------------------------------------------
# eigen.pyx surface
# ------------------------------------------------------------------------------------------------
cdef class VectorXf:
    """Same buffer contract as the reference's VectorXf (pydensecrf/eigen.pyx:11-45)."""
    cdef dcrf_mat_t* h
    cdef Py_ssize_t shape[1]
    cdef Py_ssize_t strides[1]

    def __cinit__(self):
        self.h = NULL
        self.shape[0] = 0
        self.strides[0] = 0

    def __dealloc__(self):
        if self.h != NULL:
            dcrf_mat_destroy(self.h)
            self.h = NULL

    cdef VectorXf _adopt(self, dcrf_mat_t* h):
        if self.h != NULL:
            dcrf_mat_destroy(self.h)
        self.h = h
        self.shape[0] = dcrf_mat_rows(h)
        self.strides[0] = sizeof(float)
        return self

    def __getbuffer__(self, Py_buffer* buf, int flags):
        cdef float* p = NULL
        if self.h == NULL:
            _check(dcrf_mat_create_empty(&self.h))
        _check(dcrf_mat_host_ptr(self.h, &p))
        buf.buf = <char*>p
        buf.format = 'f'
        buf.internal = NULL
        buf.itemsize = sizeof(float)
        buf.len = self.shape[0] * buf.itemsize
        buf.ndim = 1
        buf.obj = self
        buf.readonly = 0
        buf.shape = self.shape
        buf.strides = self.strides
        buf.suboffsets = NULL

    def __releasebuffer__(self, Py_buffer* buf):
        if self.h != NULL:
            dcrf_mat_release_host(self.h)


def vectorXf(float[::1] arr not None):
    cdef dcrf_mat_t* h = NULL
    cdef const float* p = &arr[0] if arr.shape[0] > 0 else NULL
    _check(dcrf_vec_from_host(p, <int>arr.shape[0], &h))
    return VectorXf()._adopt(h)


cdef class MatrixXf:
    """Handle-backed stand-in for the reference's Eigen-owning MatrixXf (pydensecrf/eigen.pyx:58-98).

    Exposes the buffer protocol with shape (rows, cols) and strides (4, 4*rows) (column-major, i.e.
    the labels of one variable are contiguous), writable.  The data may live on the GPU; the first
    buffer request downloads it."""
    cdef dcrf_mat_t* h
    cdef Py_ssize_t shape[2]
    cdef Py_ssize_t strides[2]
    cdef object owner

    def __cinit__(self):
        self.h = NULL
        self.owner = None
        self.shape[0] = self.shape[1] = 0
        self.strides[0] = self.strides[1] = 0

    def __dealloc__(self):
        if self.h != NULL:
            dcrf_mat_destroy(self.h)
            self.h = NULL

    cdef dcrf_mat_t* handle(self) except NULL:
        if self.h == NULL:
            _check(dcrf_mat_create_empty(&self.h))
        return self.h

    cdef MatrixXf _adopt(self, dcrf_mat_t* h, object owner):
        if self.h != NULL:
            dcrf_mat_destroy(self.h)
        self.h = h
        self.owner = owner
        self._refresh()
        return self

    cdef void _refresh(self):
        self.shape[0] = dcrf_mat_rows(self.h)
        self.shape[1] = dcrf_mat_cols(self.h)
        # Eigen's MatrixXf is column-major (pydensecrf/eigen.pyx:76-79)
        self.strides[0] = sizeof(float)
        self.strides[1] = self.shape[0] * self.strides[0]

    def __getbuffer__(self, Py_buffer* buf, int flags):
        cdef float* p = NULL
        cdef dcrf_mat_t* h = self.handle()
        cdef int st
        with nogil:
            st = dcrf_mat_host_ptr(h, &p)
        _check(st)
        self._refresh()
        buf.buf = <char*>p
        buf.format = 'f'
        buf.internal = NULL
        buf.itemsize = sizeof(float)
        buf.len = self.shape[0] * self.shape[1] * buf.itemsize
        buf.ndim = 2
        buf.obj = self
        buf.readonly = 0
        buf.shape = self.shape
        buf.strides = self.strides
        buf.suboffsets = NULL

    def __releasebuffer__(self, Py_buffer* buf):
        if self.h != NULL:
            dcrf_mat_release_host(self.h)

    def _argmax_device(self, bint wide=False):
        """Column-wise argmax on the GPU (int16, or int64 if `wide`; length cols), or None when the current contents
        are on the host."""
        import numpy as np
        if self.h == NULL or self.shape[1] == 0:
            return None
        # page-locked (pooled) result memory: the label map arrives by DMA at PCIe rate
        out = _result_empty((self.shape[1],), np.int64 if wide else np.int16)
        cdef size_t addr = out.ctypes.data
        cdef dcrf_mat_t* h = self.h
        cdef int st, nb = 8 if wide else 2
        with nogil:
            st = dcrf_mat_argmax(h, <void*>addr, nb)
        if st == 1:      # DCRF_EINVAL: no current device copy
            return None
        _check(st)
        return out

    def __array_function__(self, func, types, args, kwargs):
        """NumPy dispatch (NEP 18).  `np.argmax(Q, axis=0)` -- the line every caller of the reference writes after
        inference (examples/inference.py:122, README.md:199) -- is answered by the device-side argmax when Q still lives
        on the GPU: 2 bytes per variable cross PCIe instead of 4*nlabels, and no host-side reduction.  Every other NumPy
        function sees the matrix through its buffer view exactly as before."""
        import numpy as np
        if func is np.argmax and len(args) >= 1 and args[0] is self and len(args) <= 2 \
                and kwargs.get("out") is None and not kwargs.get("keepdims", False):
            axis = args[1] if len(args) == 2 else kwargs.get("axis", None)
            if axis is not None and not isinstance(axis, bool) and int(axis) in (0, -2) and self.shape[0] > 0:
                lab = self._argmax_device(wide=True)     # int64 straight from the device: no host-side conversion pass
                if lab is not None:
                    return lab if lab.dtype == np.intp else lab.astype(np.intp)
        impl = getattr(func, "_implementation", None)
        if impl is None:
            return NotImplemented
        conv = lambda a: np.asarray(a) if isinstance(a, (MatrixXf, VectorXf)) else a
        return impl(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})

    @property
    def __cuda_array_interface__(self):
        """Zero-copy device view for GPU-resident consumers (torch.as_tensor(Q, device="cuda"), cupy.asarray(Q)):
        shape (cols, rows) = (variables, labels), row pitch padded to a multiple of four floats.  Only available
        while the current contents live on the GPU (i.e. for results of inference / stepInference)."""
        cdef float* p = NULL
        cdef int pitch = 0, dev = 0
        # hasattr() probes (cupy.asarray, numba, ...) only swallow AttributeError: a host-resident matrix has no such
        # attribute, and probing must not allocate anything
        if self.h == NULL or not dcrf_mat_on_device(self.h):
            raise AttributeError("__cuda_array_interface__: the matrix' current contents are on the host")
        _check(dcrf_mat_device_ptr(self.h, &p, &pitch, &dev))
        self._refresh()
        return {"shape": (self.shape[1], self.shape[0]), "typestr": "<f4", "data": (<size_t>p, False),
                "strides": (pitch * 4, 4), "version": 3}


def matrixXf(float[:, ::1] arr not None):
    cdef dcrf_mat_t* h = NULL
    cdef const float* p = &arr[0, 0] if arr.shape[0] > 0 and arr.shape[1] > 0 else NULL
    _check(dcrf_mat_from_host(p, <int>arr.shape[0], <int>arr.shape[1], &h))
    return MatrixXf()._adopt(h, None)


# ------------------------------------------------------------------------------------------------
# densecrf.pyx surface
# ------------------------------------------------------------------------------------------------
cdef class _Compat:
    """What _labelcomp() decides in the reference (pydensecrf/densecrf.pyx:11-20): a number is Potts,
    a 1-D float32 buffer Diagonal, a 2-D float32 buffer Matrix, anything deeper a ValueError."""
    cdef int kind
    cdef float scalar
    cdef float[::1] vec
    cdef float[:, ::1] mat

    def __cinit__(self, compat, int nlabels):
        if isinstance(compat, Number):
            self.kind = 0
            self.scalar = compat
        elif memoryview(compat).ndim == 1:
            self.kind = 1
            self.vec = compat  # raises "Buffer dtype mismatch" (ValueError) for non-float32, like the reference
            if self.vec.shape[0] != nlabels:
                raise ValueError("Bad shape for diagonal label compatibility (Need ({},), got ({},))".format(nlabels, self.vec.shape[0]))
        elif memoryview(compat).ndim == 2:
            self.kind = 2
            self.mat = compat
            if self.mat.shape[0] != nlabels or self.mat.shape[1] != nlabels:
                raise ValueError("Bad shape for matrix label compatibility (Need {}, got {})".format((nlabels, nlabels), (self.mat.shape[0], self.mat.shape[1])))
        else:
            raise ValueError("LabelCompatibility of dimension >2 not meaningful.")

    cdef const float* data(self):
        if self.kind == 0:
            return &self.scalar
        if self.kind == 1:
            return &self.vec[0]
        return &self.mat[0, 0]


cdef class Unary:
    """Base of ConstUnary / LogisticUnary (pydensecrf/densecrf.pyx:23-46).  setUnary() consumes it,
    mirroring the ownership hand-off of the reference's Unary.move()."""
    cdef object _u   # float32 C-contiguous arrays kept until consumed
    cdef object _f
    cdef bint _logistic


cdef class ConstUnary(Unary):
    def __cinit__(self, float[:, ::1] u not None):
        self._u = u.copy()
        self._f = None
        self._logistic = False


cdef class LogisticUnary(Unary):
    def __cinit__(self, float[:, ::1] L not None, float[:, ::1] f not None):
        if L.shape[1] != f.shape[0]:
            raise ValueError("Bad shapes for logistic unary: L is {}, f is {}".format((L.shape[0], L.shape[1]), (f.shape[0], f.shape[1])))
        self._u = L.copy()
        self._f = f.copy()
        self._logistic = True


def empty_cache():
    """Give the library's cached device buffers, page-locked host buffers and memory-pool reserves back to the driver."""
    dcrf_empty_cache()


def map_batch_2d(float[:, :, ::1] unaries not None, unsigned char[:, :, :, ::1] images not None, sxy_gaussian=3,
                 compat_gaussian=3, sxy_bilateral=80, srgb=13, compat_bilateral=10, int n_iter=5, int workers=0,
                 device=None, bint return_q=False):
    """MAP labelling of a batch of images of one size with shared kernel parameters (SURVEY 8f-3, BASELINE cfg 4).

    unaries: (B, nlabels, H*W) float32, images: (B, H, W, 3) uint8.  Per image this is DenseCRF2D(W, H, nlabels) +
    setUnaryEnergy + addPairwiseGaussian(sxy_gaussian, compat_gaussian) + addPairwiseBilateral(sxy_bilateral, srgb,
    image, compat_bilateral) + inference(n_iter) + argmax, with identical results, but the images are spread over
    `workers` host threads inside the library (each with its own model and CUDA stream; default 4 per GPU), so that
    lattice construction, uploads and kernels of different images overlap.  `device` may be a list of GPU indices: the
    workers are then spread over those GPUs (independent replicas, no collective).  Returns (B, H, W) int16 labels, and Q (B, nlabels, H*W) float32 as
    well if return_q.  Inputs in page-locked memory (utils.pinned_empty) upload about five times faster."""
    import numpy as np
    cdef int B = <int>unaries.shape[0], L = <int>unaries.shape[1]
    cdef int H = <int>images.shape[1], W = <int>images.shape[2]
    if images.shape[0] != B:
        raise ValueError("unaries and images disagree on the batch size ({} vs {})".format(B, images.shape[0]))
    if images.shape[3] != 3:
        raise ValueError("images must be (B, H, W, 3) uint8")
    if unaries.shape[2] != H * W:
        raise ValueError("Bad shape for unary energy (Need {}, got {})".format((B, L, H * W), tuple(unaries.shape)))
    if isinstance(sxy_gaussian, Number):
        sxy_gaussian = (sxy_gaussian, sxy_gaussian)
    if isinstance(sxy_bilateral, Number):
        sxy_bilateral = (sxy_bilateral, sxy_bilateral)
    if isinstance(srgb, Number):
        srgb = (srgb, srgb, srgb)
    if not isinstance(compat_gaussian, Number) or not isinstance(compat_bilateral, Number):
        raise ValueError("map_batch_2d takes Potts compatibilities (numbers)")
    labels = np.empty((B, H, W), dtype=np.int16)
    q = np.empty((B, L, H * W), dtype=np.float32) if return_q else None
    cdef int16_t[:, :, ::1] lab_v = labels
    cdef float[:, :, ::1] q_v = q
    cdef const float** up = <const float**>malloc(max(B, 1) * sizeof(void*))
    cdef const uint8_t** ip = <const uint8_t**>malloc(max(B, 1) * sizeof(void*))
    cdef int16_t** lp = <int16_t**>malloc(max(B, 1) * sizeof(void*))
    cdef float** qp = <float**>malloc(max(B, 1) * sizeof(void*))
    cdef int b, st
    cdef int dev = -1
    cdef int ndev = 0
    cdef int devs[64]
    cdef const int* dp = NULL
    if device is not None:
        if isinstance(device, (list, tuple)):
            if not 1 <= len(device) <= 64:
                raise ValueError("device list must name 1..64 GPUs")
            ndev = len(device)
            for b in range(ndev):
                devs[b] = int(device[b])
            dp = devs
        else:
            dev = int(device)
    cdef float gsx = sxy_gaussian[0], gsy = sxy_gaussian[1], gw = compat_gaussian
    cdef float bsx = sxy_bilateral[0], bsy = sxy_bilateral[1], sr = srgb[0], sg = srgb[1], sb = srgb[2], bw = compat_bilateral
    try:
        if up == NULL or ip == NULL or lp == NULL or qp == NULL:
            raise MemoryError()
        for b in range(B):
            up[b] = &unaries[b, 0, 0]
            ip[b] = &images[b, 0, 0, 0]
            lp[b] = &lab_v[b, 0, 0]
            qp[b] = &q_v[b, 0, 0] if return_q else NULL
        with nogil:
            st = dcrf_batch_map_2d(W, H, L, dev, B, up, ip, gsx, gsy, gw, bsx, bsy, sr, sg, sb, bw, n_iter, workers,
                                   lp, qp if return_q else NULL, dp, ndev)
        _check(st)
    finally:
        free(up)
        free(ip)
        free(lp)
        free(qp)
    return (labels, q) if return_q else labels


cdef int _device_arg(dict kw) except? -2:
    dev = kw.get("device", None)
    return -1 if dev is None else int(dev)


cdef class DenseCRF:
    cdef dcrf_t* _h
    cdef int _nlabel
    cdef int _nvar

    def __cinit__(self, int nvar, int nlabels, *_, **kw):
        # like the reference, the base __cinit__ also runs for DenseCRF2D (with w, h as its first two
        # arguments) and must then leave construction to the subclass (pydensecrf/densecrf.pyx:51-63)
        self._h = NULL
        if not isinstance(self, DenseCRF2D):
            _check(dcrf_create(nvar, nlabels, _device_arg(kw), &self._h))
        self._nvar = nvar
        self._nlabel = nlabels

    def __dealloc__(self):
        if self._h != NULL:
            dcrf_destroy(self._h)
            self._h = NULL

    def addPairwiseEnergy(self, float[:, ::1] features not None, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if features.shape[1] != self._nvar:
            raise ValueError("Bad shape for pairwise energy (Need (?, {}), got {})".format(self._nvar, (features.shape[0], features.shape[1])))
        cdef _Compat c = _Compat(compat, self._nlabel)
        cdef int kt = int(KernelType(kernel)), nt = int(NormalizationType(normalization))
        cdef int d = <int>features.shape[0], st
        if d < 1:
            raise ValueError("Bad shape for pairwise energy (need at least one feature row)")
        cdef const float* fp = &features[0, 0]
        cdef const float* cp = c.data()
        with nogil:
            st = dcrf_add_pairwise_energy(self._h, fp, d, c.kind, cp, kt, nt)
        _check(st)

    def setUnary(self, Unary u not None):
        cdef float[:, ::1] a, b
        cdef int st
        if u._u is None:
            raise ValueError("This Unary object was already handed to a DenseCRF (setUnary takes ownership).")
        a = u._u
        if u._logistic:
            b = u._f
            if a.shape[0] != self._nlabel or b.shape[1] != self._nvar:
                raise ValueError("Bad shape for logistic unary (Need L ({}, F) and f (F, {}), got {} and {})".format(
                    self._nlabel, self._nvar, (a.shape[0], a.shape[1]), (b.shape[0], b.shape[1])))
            with nogil:
                st = dcrf_set_unary_logistic(self._h, &a[0, 0], <int>a.shape[1], &b[0, 0])
        else:
            if a.shape[0] != self._nlabel or a.shape[1] != self._nvar:
                raise ValueError("Bad shape for unary energy (Need {}, got {})".format((self._nlabel, self._nvar), (a.shape[0], a.shape[1])))
            with nogil:
                st = dcrf_set_unary_const(self._h, &a[0, 0])
        _check(st)
        u._u = None
        u._f = None

    def setUnaryEnergy(self, float[:, ::1] u not None, float[:, ::1] f=None):
        if u.shape[0] != self._nlabel or u.shape[1] != self._nvar:
            raise ValueError("Bad shape for unary energy (Need {}, got {})".format((self._nlabel, self._nvar), (u.shape[0], u.shape[1])))
        cdef int st
        if f is None:
            with nogil:
                st = dcrf_set_unary_const(self._h, &u[0, 0])
        else:
            # the reference forwards (u, f) to LogisticUnaryEnergy(L=u, feature=f) after the shape check
            # above (pydensecrf/densecrf.pyx:80-89), so f has to be (nvar, nvar) for L*f to be a unary
            if f.shape[0] != u.shape[1] or f.shape[1] != self._nvar:
                raise ValueError("Bad shape for logistic feature (Need {}, got {})".format((u.shape[1], self._nvar), (f.shape[0], f.shape[1])))
            with nogil:
                st = dcrf_set_unary_logistic(self._h, &u[0, 0], <int>u.shape[1], &f[0, 0])
        _check(st)

    def setUnaryFromLabels(self, labels, double gt_prob, bint zero_unsure=True):
        """setUnaryEnergy(utils.unary_from_labels(labels, nlabels, gt_prob, zero_unsure)) (utils.py:6-46) with the energies
        formed on the device: the label map is what crosses the bus.  Same unary bit for bit."""
        import numpy as np
        if not (0 < gt_prob < 1):
            raise AssertionError("`gt_prob must be in (0,1).")
        arr = np.ascontiguousarray(np.asarray(labels).flatten(), dtype=np.int32)
        if arr.shape[0] != self._nvar:
            raise ValueError("Bad shape for labels (Need {} entries, got {})".format(self._nvar, arr.shape[0]))
        cdef int32_t[::1] v = arr
        cdef int st
        with nogil:
            st = dcrf_set_unary_from_labels(self._h, &v[0], gt_prob, 1 if zero_unsure else 0)
        if st == 1 and b"out of bounds" in dcrf_last_error():
            raise IndexError(dcrf_last_error().decode())
        _check(st)

    def inference(self, int niter):
        cdef dcrf_mat_t* q = NULL
        cdef int st
        with nogil:
            st = dcrf_inference(self._h, niter, &q)
        _check(st)
        return MatrixXf()._adopt(q, self)

    def startInference(self):
        cdef dcrf_mat_t* q = NULL
        cdef int st
        with nogil:
            st = dcrf_start_inference(self._h, &q)
        _check(st)
        return MatrixXf()._adopt(q, self), MatrixXf(), MatrixXf()

    def stepInference(self, MatrixXf Q not None, MatrixXf tmp1 not None, MatrixXf tmp2 not None):
        cdef dcrf_mat_t* q = Q.handle()
        cdef dcrf_mat_t* t1 = tmp1.handle()
        cdef dcrf_mat_t* t2 = tmp2.handle()
        cdef int st
        with nogil:
            st = dcrf_step_inference(self._h, q, t1, t2)
        _check(st)

    def klDivergence(self, MatrixXf Q not None):
        cdef dcrf_mat_t* q = Q.handle()
        cdef double out = 0.0
        cdef int st
        with nogil:
            st = dcrf_kl_divergence(self._h, q, &out)
        _check(st)
        return out

    # ---- additions (not in the reference's wrapper) ------------------------------------------------
    def setUnaryEnergyDevice(self, u):
        """setUnaryEnergy for a unary that already lives in GPU memory: `u` is any object exposing
        __cuda_array_interface__ (torch / cupy array), float32, C-contiguous, shape (nlabels, nvar), on the model's
        device.  No host round trip."""
        cai = u.__cuda_array_interface__
        shape = tuple(cai["shape"])
        if len(shape) != 2 or shape[0] != self._nlabel or shape[1] != self._nvar:
            raise ValueError("Bad shape for unary energy (Need {}, got {})".format((self._nlabel, self._nvar), shape))
        if cai["typestr"] not in ("<f4", "=f4", "|f4"):
            raise ValueError("Buffer dtype mismatch, expected 'float' but got {}".format(cai["typestr"]))
        st_ = cai.get("strides")
        if st_ is not None and tuple(st_) != (shape[1] * 4, 4):
            raise ValueError("device array is not C-contiguous")
        cdef size_t ptr = cai["data"][0]
        cdef int st
        with nogil:
            st = dcrf_set_unary_device(self._h, <const float*>ptr)
        _check(st)

    def setUnaryFromSoftmax(self, sm, scale=None, clip=1e-5):
        """utils.unary_from_softmax(sm, scale, clip) + setUnaryEnergy in one call, evaluated on the GPU while the
        probabilities are uploaded (SURVEY 8f-2).  `sm`: class probabilities, classes on axis 0."""
        import numpy as np
        a = np.ascontiguousarray(np.asarray(sm, dtype=np.float32).reshape(sm.shape[0], -1))
        if a.shape[0] != self._nlabel or a.shape[1] != self._nvar:
            raise ValueError("Bad shape for softmax (Need {}, got {})".format((self._nlabel, self._nvar), (a.shape[0], a.shape[1])))
        if scale is not None and not (0 < scale <= 1):
            raise ValueError("`scale` needs to be in (0,1]")
        cdef float[:, ::1] v = a
        cdef double sc = -1.0 if scale is None else scale
        cdef double cl = -1.0 if clip is None else clip
        cdef int st
        with nogil:
            st = dcrf_set_unary_from_softmax(self._h, &v[0, 0], sc, cl)
        _check(st)

    def _add_nd(self, sdims, schan, img, chdim, compat, kernel, normalization, shape):
        import numpy as np
        cdef _Compat c = _Compat(compat, self._nlabel)
        cdef int kt = int(KernelType(kernel)), nt = int(NormalizationType(normalization)), st
        nch = 0
        flat = None
        if img is not None:
            im = np.asarray(img)
            if chdim == -1:
                im = im[..., np.newaxis]
            else:
                im = np.moveaxis(im, chdim, -1)
            nd = im.ndim                 # (no negative indices: this module is compiled with wraparound=False)
            shape = im.shape[:nd - 1]
            nch = im.shape[nd - 1]
            if im.dtype != np.uint8:
                im = im.astype(np.float32)
            flat = np.ascontiguousarray(im.reshape(-1, nch))
            if isinstance(schan, Number):
                schan = [schan] * nch
            if len(schan) != nch:
                raise ValueError("need one `schan` entry per channel ({}), got {}".format(nch, len(schan)))
        shape = tuple([int(x) for x in shape])
        if len(sdims) != len(shape):
            raise ValueError("need one `sdims` entry per spatial axis ({}), got {}".format(len(shape), len(sdims)))
        if int(np.prod(shape)) != self._nvar:
            raise ValueError("Bad shape for pairwise energy (shape {} has {} elements, need {})".format(shape, int(np.prod(shape)), self._nvar))
        d_arr = np.ascontiguousarray(shape, dtype=np.int32)
        s_arr = np.ascontiguousarray(sdims, dtype=np.float32)
        c_arr = np.ascontiguousarray(schan if nch else [0.0], dtype=np.float32)
        cdef int[::1] dv = d_arr
        cdef float[::1] sv = s_arr
        cdef float[::1] cv = c_arr
        cdef const float* cp = c.data()
        cdef int nsp = len(shape), nc = nch, isf = 0
        cdef const void* ip = NULL
        cdef unsigned char[:, ::1] u8
        cdef float[:, ::1] f32
        if nch:
            if flat.dtype == np.uint8:
                u8 = flat
                ip = <const void*>&u8[0, 0]
            else:
                f32 = flat
                ip = <const void*>&f32[0, 0]
                isf = 1
        with nogil:
            st = dcrf_add_pairwise_nd(self._h, nsp, &dv[0], &sv[0], nc, &cv[0], ip, isf, c.kind, cp, kt, nt)
        _check(st)

    def addPairwiseGaussianND(self, sdims, shape, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        """addPairwiseEnergy(utils.create_pairwise_gaussian(sdims, shape), ...) with the features generated on the GPU."""
        self._add_nd(sdims, None, None, -1, compat, kernel, normalization, shape)

    def addPairwiseBilateralND(self, sdims, schan, img, compat, chdim=-1, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        """addPairwiseEnergy(utils.create_pairwise_bilateral(sdims, schan, img, chdim), ...) with the features generated on
        the GPU: only the image (uint8 or float32) is uploaded instead of the (n_spatial + n_channels) x N float matrix."""
        self._add_nd(sdims, schan, img, chdim, compat, kernel, normalization, None)

    def inferenceInto(self, int niter, MatrixXf Q not None):
        """inference(niter) written into an existing MatrixXf (no allocation when shapes match)."""
        cdef dcrf_mat_t* q = Q.handle()
        cdef int st
        with nogil:
            st = dcrf_inference_into(self._h, niter, q)
        _check(st)
        Q.owner = self
        Q._refresh()
        return Q

    def currentMap(self, MatrixXf Q not None):
        """Per-variable argmax of Q computed on the device (DenseCRF::currentMap, densecrf.cpp:202-211):
        int16 array of length nvar; only 2 bytes per variable cross PCIe instead of the 4*nlabels of Q."""
        import numpy as np
        out = _result_empty((self._nvar,), np.int16)
        cdef int16_t[::1] o = out
        cdef dcrf_mat_t* q = Q.handle()
        cdef int st
        with nogil:
            st = dcrf_current_map(self._h, q, &o[0])
        _check(st)
        return out

    def map(self, int n_iterations):
        """DenseCRF::map (densecrf.cpp:132-137): inference(n_iterations) followed by currentMap; Q never leaves the GPU.
        int16 array of length nvar."""
        import numpy as np
        out = _result_empty((self._nvar,), np.int16)
        cdef int16_t[::1] o = out
        cdef int st
        with nogil:
            st = dcrf_map(self._h, n_iterations, &o[0])
        _check(st)
        return out

    def synchronize(self):
        cdef int st
        with nogil:  # other Python threads keep driving their own models while this one waits
            st = dcrf_synchronize(self._h)
        _check(st)

    def _timer_start(self):
        _check(dcrf_timer_start(self._h))

    def _timer_stop(self):
        cdef float ms = 0
        cdef int st
        with nogil:
            st = dcrf_timer_stop(self._h, &ms)
        _check(st)
        return ms

    def _profile_enable(self, bint on):
        _check(dcrf_profile_enable(self._h, 1 if on else 0))

    def _profile_read(self, bint reset=True):
        cdef dcrf_prof_entry ent[64]
        cdef int n = 0
        _check(dcrf_profile_read(self._h, ent, 64, &n, 1 if reset else 0))
        out = {}
        for i in range(min(n, 64)):
            out[ent[i].name.decode()] = (ent[i].launches, ent[i].ms_total)
        return out

    def _launch_count(self):
        return dcrf_launch_count(self._h)

    def _flush_l2(self):
        _check(dcrf_flush_l2(self._h))

    def _selftest(self, int which, n=0, seed=0):
        """Device self-test of the restated arithmetic (0: packed pexp, exhaustive; 1: column division, n random
        pairs); returns the number of mismatching results."""
        cdef uint64_t bad = 0, nn = n, sd = seed
        cdef int st
        with nogil:
            st = dcrf_debug_selftest(self._h, which, nn, sd, &bad)
        _check(st)
        return int(bad)

    def _set_iteration_mode(self, int mode):
        _check(dcrf_set_iteration_mode(self._h, mode))

    def setSoftmaxMode(self, bint fast):
        """False (default): bit-faithful softmax; True: faster approximate softmax (~1e-6 relative)."""
        _check(dcrf_set_softmax_mode(self._h, 1 if fast else 0))

    def setKLMode(self, int mode):
        """0 (default): klDivergence sums in double through a fixed tree (deterministic, ~1e-4 relative of the reference,
        whose float reduction is the less accurate side); 1: the reference's float reduction order restated (pairwise terms
        bit-exact, slower)."""
        _check(dcrf_set_kl_mode(self._h, mode))

    def _set_kernel_variant(self, int mask):
        """Bit mask: 1 wide-lane update kernel, 2 pre-scaled Q copies for the splats.  Identical results; for A/B timing."""
        _check(dcrf_set_kernel_variant(self._h, mask))

    def setMatrixCompatMode(self, int mode):
        """0 (default): tensor cores (3xTF32) for MatrixCompatibility from 64 labels up; 1: always the
        bit-faithful exact-order kernel; 2: tensor cores whenever the label count allows (16..256)."""
        _check(dcrf_set_matrix_mode(self._h, mode))

    def _num_pairwise(self):
        return dcrf_num_pairwise(self._h)

    def _lattice_info(self, int k):
        cdef int d = 0, N = 0, V = 0
        _check(dcrf_lattice_info(self._h, k, &d, &N, &V))
        return d, N, V

    def _lattice_dump(self, int k):
        """(offset (N,d+1) int32, bary (N,d+1) f32, keys (V,d) int16, nbr (d+1,V,2) int32, norm (N,) f32)"""
        import numpy as np
        d, N, V = self._lattice_info(k)
        off = np.zeros((N, d + 1), np.int32)
        bary = np.zeros((N, d + 1), np.float32)
        keys = np.zeros((max(V, 1), d), np.int16)
        nbr = np.zeros((d + 1, max(V, 1), 2), np.int32)
        norm = np.zeros(N, np.float32)
        cdef int32_t[:, ::1] o = off
        cdef float[:, ::1] b = bary
        cdef int16_t[:, ::1] ky = keys
        cdef int32_t[:, :, ::1] nb = nbr
        cdef float[::1] nm = norm
        cdef int st
        with nogil:
            st = dcrf_lattice_dump(self._h, k, &o[0, 0], &b[0, 0], &ky[0, 0], &nb[0, 0, 0], &nm[0])
        _check(st)
        return off, bary, keys[:V], nbr[:, :V], norm

    def _lattice_filter(self, int k, float[:, ::1] x_NxL not None, bint reverse=False):
        """Permutohedral::compute of term k on a (N, L) pixel-major array."""
        import numpy as np
        out = np.empty((x_NxL.shape[0], x_NxL.shape[1]), np.float32)
        cdef float[:, ::1] o = out
        cdef int L = <int>x_NxL.shape[1], st
        cdef int rev = 1 if reverse else 0
        d, N, V = self._lattice_info(k)
        if x_NxL.shape[0] != N:
            raise ValueError("need (N, L) input with N = {}".format(N))
        with nogil:
            st = dcrf_lattice_filter(self._h, k, &x_NxL[0, 0], L, rev, &o[0, 0])
        _check(st)
        return out

    def _pairwise_apply(self, int k, MatrixXf Q not None, bint transpose=False):
        """PairwisePotential::apply / applyTranspose of term k (pairwise.cpp:173-183)."""
        cdef dcrf_mat_t* o = NULL
        cdef dcrf_mat_t* q = Q.handle()
        cdef int st
        cdef bint tr = transpose
        with nogil:
            st = dcrf_pairwise_apply_transpose(self._h, k, q, &o) if tr else dcrf_pairwise_apply(self._h, k, q, &o)
        _check(st)
        return MatrixXf()._adopt(o, self)

    def _lattice_gradient(self, int k, float[:, ::1] a_NxL not None, float[:, ::1] b_NxL not None):
        """Permutohedral::gradient of term k's lattice (permutohedral.cpp:612-695): (N, d) float32, d(a^T K b)/d features."""
        import numpy as np
        d, N, V = self._lattice_info(k)
        if a_NxL.shape[0] != N or b_NxL.shape[0] != N or a_NxL.shape[1] != b_NxL.shape[1]:
            raise ValueError("need two (N, L) inputs with N = {}".format(N))
        out = np.empty((N, d), np.float32)
        cdef float[:, ::1] o = out
        cdef int L = <int>a_NxL.shape[1], st
        with nogil:
            st = dcrf_lattice_gradient(self._h, k, &a_NxL[0, 0], &b_NxL[0, 0], L, &o[0, 0])
        _check(st)
        return out


    def _labels_arg(self, labels):
        import numpy as np
        arr = np.ascontiguousarray(np.asarray(labels).flatten(), dtype=np.int16)
        if arr.shape[0] != self._nvar:
            raise ValueError("Bad shape for labels (Need {} entries, got {})".format(self._nvar, arr.shape[0]))
        return arr

    def unaryEnergy(self, labels):
        """DenseCRF::unaryEnergy (densecrf.cpp:141-153): float32 (nvar,), unary(l_i, i) of every variable."""
        import numpy as np
        arr = self._labels_arg(labels)
        out = np.empty(self._nvar, np.float32)
        cdef int16_t[::1] v = arr
        cdef float[::1] o = out
        cdef int st
        with nogil:
            st = dcrf_unary_energy(self._h, &v[0], &o[0])
        _check(st)
        return out

    def pairwiseEnergy(self, labels, int term=-1):
        """DenseCRF::pairwiseEnergy (densecrf.cpp:154-177): float32 (nvar,), -0.5 * apply(one-hot belief)(l_i, i) of pairwise
        term `term`, or of all terms summed (term = -1)."""
        import numpy as np
        arr = self._labels_arg(labels)
        out = np.empty(self._nvar, np.float32)
        cdef int16_t[::1] v = arr
        cdef float[::1] o = out
        cdef int st
        with nogil:
            st = dcrf_pairwise_energy(self._h, &v[0], term, &o[0])
        _check(st)
        return out

    # ---- learning path (densecrf.h:106-124; not wrapped by the reference's densecrf.pyx) ------------------------------
    def _get_params(self, int group):
        import numpy as np
        cdef int n = 0
        _check(dcrf_num_parameters(self._h, group, &n))
        out = np.zeros(n, np.float32)
        cdef float[::1] o = out
        if n > 0:
            _check(dcrf_get_parameters(self._h, group, &o[0]))
        return out

    def _set_params(self, int group, v):
        import numpy as np
        arr = np.ascontiguousarray(v, dtype=np.float32).ravel()
        cdef float[::1] a = arr
        cdef int n = <int>arr.shape[0], st
        cdef const float* p = &a[0] if n > 0 else NULL
        with nogil:
            st = dcrf_set_parameters(self._h, group, p, n)
        _check(st)

    def unaryParameters(self):
        """DenseCRF::unaryParameters (densecrf.cpp:298-302)."""
        return self._get_params(0)

    def setUnaryParameters(self, v):
        self._set_params(0, v)

    def labelCompatibilityParameters(self):
        """DenseCRF::labelCompatibilityParameters (densecrf.cpp:307-320)."""
        return self._get_params(1)

    def setLabelCompatibilityParameters(self, v):
        self._set_params(1, v)

    def kernelParameters(self):
        """DenseCRF::kernelParameters (densecrf.cpp:335-348)."""
        return self._get_params(2)

    def setKernelParameters(self, v):
        """DenseCRF::setKernelParameters (densecrf.cpp:349-363): every DIAG / FULL kernel rebuilds its lattice."""
        self._set_params(2, v)

    def gradient(self, int n_iterations, objective, bint unary=True, bint label_compatibility=True, bint kernel=True):
        """DenseCRF::gradient (densecrf.cpp:237-297): (objective value, unary_grad, lbl_cmp_grad, kernel_grad) after
        n_iterations of mean field; a gradient that was not asked for (or whose parameter vector is empty) comes back
        as an empty array.  `objective` is a LogLikelihood, Hamming or IntersectionOverUnion."""
        import numpy as np
        cdef Objective ob = objective
        cdef dcrf_objective o
        cdef int16_t[::1] gt = ob.gt
        if gt.shape[0] != self._nvar:
            raise ValueError("the objective's ground truth has {} entries, the model {} variables".format(gt.shape[0], self._nvar))
        cdef float[::1] cw = ob.class_weight
        o.kind = ob.kind
        o.gt = &gt[0]
        o.robust = ob.robust
        o.class_weight = &cw[0] if cw.shape[0] > 0 else NULL
        o.n_class_weight = <int>cw.shape[0]
        gu = np.zeros(len(self._get_params(0)) if unary else 0, np.float32)
        gl = np.zeros(len(self._get_params(1)) if label_compatibility else 0, np.float32)
        gk = np.zeros(len(self._get_params(2)) if kernel else 0, np.float32)
        cdef float[::1] vu = gu, vl = gl, vk = gk
        cdef float* pu = &vu[0] if gu.shape[0] > 0 else NULL
        cdef float* pl = &vl[0] if gl.shape[0] > 0 else NULL
        cdef float* pk = &vk[0] if gk.shape[0] > 0 else NULL
        cdef double value = 0
        cdef int st
        with nogil:
            st = dcrf_gradient(self._h, n_iterations, &o, &value, pu, pl, pk)
        _check(st)
        return value, gu, gl, gk


cdef class Objective:
    """ObjectiveFunction (objective.h:31-36): what DenseCRF.gradient evaluates on the last Q."""
    cdef int kind
    cdef float robust
    cdef object gt            # int16 (N,)
    cdef object class_weight  # float32

    def _init(self, int kind, gt, float robust, w):
        import numpy as np
        self.kind = kind
        self.gt = np.ascontiguousarray(gt, dtype=np.int16).ravel()
        self.robust = robust
        self.class_weight = np.ascontiguousarray(w, dtype=np.float32).ravel()


cdef class LogLikelihood(Objective):
    """LogLikelihood(gt, robust=0) (objective.h:38-46, objective.cpp:35-50)."""
    def __init__(self, gt, float robust=0):
        self._init(0, gt, robust, [])


cdef class Hamming(Objective):
    """Hamming(gt, class_weight_pow=0) or Hamming(gt, class_weights) (objective.h:48-57, objective.cpp:51-80)."""
    def __init__(self, gt, class_weight=0.0):
        import numpy as np
        g = np.ascontiguousarray(gt, dtype=np.int16).ravel()
        if np.ndim(class_weight) == 0:
            # objective.cpp:51-63, in the reference's float arithmetic
            M = max(int(g.max()) + 1, 0) if g.size else 0
            cnt = np.bincount(g[g >= 0], minlength=M).astype(np.float32)
            w = cnt / cnt.sum(dtype=np.float32)
            w = np.power(w, np.float32(-float(class_weight)), dtype=np.float32)
            w = w / (cnt * w).sum(dtype=np.float32)
            class_weight = w
        self._init(1, g, 0, class_weight)

    @property
    def class_weights(self):
        return self.class_weight


cdef class IntersectionOverUnion(Objective):
    """IntersectionOverUnion(gt) (objective.h:59-66, objective.cpp:81-109)."""
    def __init__(self, gt):
        self._init(2, gt, 0, [])


cdef class DenseCRF2D(DenseCRF):
    cdef int _w
    cdef int _h2

    def __cinit__(self, int w, int h, int nlabels, *_, **kw):
        _check(dcrf_create_2d(w, h, nlabels, _device_arg(kw), &self._h))
        self._w = w
        self._h2 = h
        self._nvar = w * h
        self._nlabel = nlabels

    def addPairwiseGaussian(self, sxy, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if isinstance(sxy, Number):
            sxy = (sxy, sxy)
        cdef _Compat c = _Compat(compat, self._nlabel)
        cdef int kt = int(KernelType(kernel)), nt = int(NormalizationType(normalization)), st
        cdef float sx = sxy[0], sy = sxy[1]
        cdef const float* cp = c.data()
        with nogil:
            st = dcrf_add_pairwise_gaussian(self._h, sx, sy, c.kind, cp, kt, nt)
        _check(st)

    def addPairwiseBilateral(self, sxy, srgb, unsigned char[:, :, ::1] rgbim not None, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if isinstance(sxy, Number):
            sxy = (sxy, sxy)
        if isinstance(srgb, Number):
            srgb = (srgb, srgb, srgb)
        if rgbim.shape[0] != self._h2 or rgbim.shape[1] != self._w:
            raise ValueError("Bad shape for pairwise bilateral (Need {}, got {})".format((self._h2, self._w, 3), (rgbim.shape[0], rgbim.shape[1], rgbim.shape[2])))
        if rgbim.shape[2] != 3:
            raise ValueError("addPairwiseBilateral only works for RGB images. For other types, use `utils.create_pairwise_bilateral` to construct your own pairwise energy and add it through `addPairwiseEnergy`.")
        cdef _Compat c = _Compat(compat, self._nlabel)
        cdef int kt = int(KernelType(kernel)), nt = int(NormalizationType(normalization)), st
        cdef float sx = sxy[0], sy = sxy[1], sr = srgb[0], sg = srgb[1], sb = srgb[2]
        cdef const float* cp = c.data()
        cdef const uint8_t* ip = &rgbim[0, 0, 0]
        with nogil:
            st = dcrf_add_pairwise_bilateral(self._h, sx, sy, sr, sg, sb, ip, c.kind, cp, kt, nt)
        _check(st)
