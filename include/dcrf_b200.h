/*
 * dcrf_b200.h -- C-ABI of libdcrf_b200.so, the B200 (sm_100a) mean-field engine behind
 * pydensecrf's Python surface.
 *
 * The reference has no C interface: its Cython layer binds the C++ classes directly
 * (pydensecrf/densecrf.pxd:4-84).  This header is the boundary the build introduces instead; each
 * entry point replaces exactly one call the reference's .pyx makes into C++ and names it below as
 * "replaces <file>:<line>".  Paths are relative to /root/reference/.  INTEGRATION.md shows the
 * .pyx/.pxd a maintainer of the reference would write against this header.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch/numpy types.
 *   - every function returns a dcrf_status; on failure dcrf_last_error() (thread-local) describes it.
 *     DCRF_EINVAL -> ValueError, DCRF_ECUDA -> RuntimeError, DCRF_ENOMEM -> MemoryError.
 *   - host arrays are read during the call only (copied / uploaded before return), like the
 *     reference, which copies every numpy input at its boundary (pydensecrf/eigen_impl.cpp:19-25).
 *   - "LxN row-major" = numpy (L, N) C-order;  "NxL pixel-major" = the memory of an
 *     Eigen::MatrixXf(L, N), which is what pydensecrf's MatrixXf exposes through the buffer
 *     protocol (pydensecrf/eigen.pyx:58-98: shape (L,N), strides (4, 4L)).
 *   - a handle is not thread-safe (same contract as the reference); distinct handles are
 *     independent and may live on different GPUs.  All work of a handle runs on its own stream.
 *   - there is no CPU fallback: any entry that computes needs a CUDA device and fails with
 *     DCRF_ECUDA otherwise.
 */
#ifndef DCRF_B200_H
#define DCRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCRF_ABI_VERSION 1

typedef enum { DCRF_OK = 0, DCRF_EINVAL = 1, DCRF_ECUDA = 2, DCRF_ENOMEM = 3 } dcrf_status;

/* pydensecrf/densecrf/include/pairwise.h:32-42 (same numeric values) */
typedef enum { DCRF_NO_NORMALIZATION = 0, DCRF_NORMALIZE_BEFORE = 1, DCRF_NORMALIZE_AFTER = 2, DCRF_NORMALIZE_SYMMETRIC = 3 } dcrf_norm_type;
typedef enum { DCRF_CONST_KERNEL = 0, DCRF_DIAG_KERNEL = 1, DCRF_FULL_KERNEL = 2 } dcrf_kernel_type;
/* which LabelCompatibility the .pyx's _labelcomp() would have built (pydensecrf/densecrf.pyx:11-20) */
typedef enum {
  DCRF_COMPAT_POTTS = 0,   /* compat_data[1]:   out = -w * Q          labelcompatibility.cpp:46-48 */
  DCRF_COMPAT_DIAGONAL = 1,/* compat_data[L]:   out = diag(w) * Q     labelcompatibility.cpp:66-69 */
  DCRF_COMPAT_MATRIX = 2   /* compat_data[L*L] row-major: out = 0.5(m+m^T) * Q   :79-84 */
} dcrf_compat_kind;

typedef struct dcrf_model dcrf_t;   /* a DenseCRF / DenseCRF2D object */
typedef struct dcrf_mat dcrf_mat_t; /* an eigen.MatrixXf / VectorXf object: host mirror + device buffer */

const char* dcrf_last_error(void);
int dcrf_abi_version(void);
/* number of visible CUDA devices; 0 when there is no driver/GPU (never fails) */
int dcrf_device_count(void);

/* ---- model lifetime --------------------------------------------------------------------------- */
/* replaces `new c_DenseCRF(nvar, nlabels)`            pydensecrf/densecrf.pyx:58  (densecrf.cpp:39)
 * device < 0 selects $DCRF_B200_DEVICE, else device 0. */
dcrf_status dcrf_create(int nvar, int nlabels, int device, dcrf_t** out);
/* replaces `new c_DenseCRF2D(w, h, nlabels)`          pydensecrf/densecrf.pyx:110 (densecrf.cpp:47) */
dcrf_status dcrf_create_2d(int width, int height, int nlabels, int device, dcrf_t** out);
/* replaces `del self._this`                           pydensecrf/densecrf.pyx:68 */
void dcrf_destroy(dcrf_t* m);

/* ---- unary ------------------------------------------------------------------------------------ */
/* replaces setUnaryEnergy(c_matrixXf(u))              pydensecrf/densecrf.pyx:87  (densecrf.cpp:89; unary.cpp:42-46)
 * and `new ConstUnaryEnergy` + setUnaryEnergy(u.move()) pydensecrf/densecrf.pyx:41,78.   u: LxN row-major. */
dcrf_status dcrf_set_unary_const(dcrf_t* m, const float* u_LxN);
/* replaces setUnaryEnergy(c_matrixXf(L), c_matrixXf(f)) pydensecrf/densecrf.pyx:89 and LogisticUnary :46,78
 * (unary.cpp:48-52: U = Lmat * f).   Lmat: L x F row-major, f: F x N row-major. */
dcrf_status dcrf_set_unary_logistic(dcrf_t* m, const float* Lmat_LxF, int F, const float* f_FxN);

/* GPU-resident pipelines: the unary already sits in device memory (e.g. a network's output).  dev_u_LxN: LxN
 * row-major floats on the model's GPU; no PCIe transfer.  Synchronises the device first (the producer's stream is
 * not known to this library). */
dcrf_status dcrf_set_unary_device(dcrf_t* m, const float* dev_u_LxN);
/* pydensecrf.utils.unary_from_softmax (utils.py:55-79) fused into the upload: U = -log(clip(scale*p +
 * (1-scale)/L, clip, 1)).  sm: LxN row-major class probabilities; scale <= 0 means "no scale", clip <= 0 "no clip".
 * NumPy's dtype rules are mirrored: with a scale the float32 product scale*p is widened and everything after it is
 * float64 (the uniform term is a float64 array), without one the clip bounds are float32; the logarithm is taken in
 * double and rounded once (numpy's float32 log may differ from that by one ulp). */
dcrf_status dcrf_set_unary_from_softmax(dcrf_t* m, const float* sm_LxN, double scale, double clip);
/* utils.unary_from_labels(labels, nlabels, gt_prob, zero_unsure) + setUnaryEnergy (pydensecrf/utils.py:6-46,
 * examples/inference.py:70-75) with the energies formed on the device: one int32 label per variable is uploaded instead of
 * nlabels floats (8 MB instead of 174 MB at 1080p / 21 labels).  Bit-identical to the two reference calls. */
dcrf_status dcrf_set_unary_from_labels(dcrf_t* m, const int32_t* labels_N, double gt_prob, int zero_unsure);

/* ---- pairwise terms (each builds one permutohedral lattice on the GPU) ------------------------- */
/* replaces addPairwiseEnergy(c_matrixXf(features), _labelcomp(compat), kernel, normalization)
 *                                                      pydensecrf/densecrf.pyx:75 (densecrf.cpp:54; pairwise.cpp:116-122,40-62;
 *                                                      permutohedral.cpp:140-321).   feat: d x N row-major. */
dcrf_status dcrf_add_pairwise_energy(dcrf_t* m, const float* feat_dxN, int d, int compat_kind,
                                     const float* compat_data, int kernel_type, int norm_type);
/* replaces _this2d.addPairwiseGaussian(sx, sy, ...)   pydensecrf/densecrf.pyx:124 (densecrf.cpp:61-69) */
dcrf_status dcrf_add_pairwise_gaussian(dcrf_t* m, float sx, float sy, int compat_kind, const float* compat_data,
                                       int kernel_type, int norm_type);
/* replaces _this2d.addPairwiseBilateral(sx,sy,sr,sg,sb,&rgbim[0,0,0], ...)  pydensecrf/densecrf.pyx:138 (densecrf.cpp:70-81)
 * rgb: H x W x 3 bytes, channel order as given. */
dcrf_status dcrf_add_pairwise_bilateral(dcrf_t* m, float sx, float sy, float sr, float sg, float sb,
                                        const uint8_t* rgb_HxWx3, int compat_kind, const float* compat_data,
                                        int kernel_type, int norm_type);

/* pydensecrf.utils.create_pairwise_gaussian / create_pairwise_bilateral (utils.py:89-162) + addPairwiseEnergy
 * with the features generated on the device: nsp spatial coordinates (C order over dims[], prod(dims) = nvar) divided
 * by sdims[], followed by nch channels of a channel-last image (N x nch, uint8 or float32) divided by schan[].
 * nch = 0 gives the Gaussian (position-only) kernel.  Bit-identical to feeding the numpy-built features. */
dcrf_status dcrf_add_pairwise_nd(dcrf_t* m, int nsp, const int* dims, const float* sdims, int nch, const float* schan,
                                 const void* img, int img_is_float, int compat_kind, const float* compat_data,
                                 int kernel_type, int norm_type);

/* ---- inference -------------------------------------------------------------------------------- */
/* replaces eigen.MatrixXf().wrap(_this.inference(niter))   pydensecrf/densecrf.pyx:92 (densecrf.cpp:115-131)
 * *Q is a new matrix (L x N) whose data stays on the device until its host view is requested. */
dcrf_status dcrf_inference(dcrf_t* m, int n_iterations, dcrf_mat_t** Q);
/* replaces _this.startInference()                     pydensecrf/densecrf.pyx:95 (densecrf.cpp:178-186) */
dcrf_status dcrf_start_inference(dcrf_t* m, dcrf_mat_t** Q);
/* replaces _this.stepInference(Q.m, tmp1.m, tmp2.m)   pydensecrf/densecrf.pyx:98 (densecrf.cpp:187-201)
 * Q is updated in place.  tmp1/tmp2 are accepted for signature compatibility and may be NULL; the
 * engine keeps its own scratch, so they are left untouched. */
dcrf_status dcrf_step_inference(dcrf_t* m, dcrf_mat_t* Q, dcrf_mat_t* tmp1, dcrf_mat_t* tmp2);
/* replaces _this.klDivergence(Q.m)                    pydensecrf/densecrf.pyx:101 (densecrf.cpp:214-235) */
dcrf_status dcrf_kl_divergence(dcrf_t* m, dcrf_mat_t* Q, double* out);
/* DenseCRF::currentMap (densecrf.cpp:202-211; unwrapped in the reference's .pyx:61-62):
 * per-pixel argmax of Q computed on the device, first maximum wins; labels_N receives N int16. */
dcrf_status dcrf_current_map(dcrf_t* m, dcrf_mat_t* Q, int16_t* labels_N);
/* DenseCRF::map(n_iterations) (densecrf.cpp:132-137; not wrapped by densecrf.pyx): inference + currentMap in one call;
 * Q stays on the device, only the int16 label map crosses PCIe. */
dcrf_status dcrf_map(dcrf_t* m, int n_iterations, int16_t* labels_N);

/* a batch of independent images of one size with shared kernel parameters (BASELINE cfg 4).  For
 * every image b the call does what a caller of the reference does per image --
 *   d = DenseCRF2D(W, H, L); d.setUnaryEnergy(unary[b]); d.addPairwiseGaussian((g_sx, g_sy), g_w);
 *   d.addPairwiseBilateral((b_sx, b_sy), (b_sr, b_sg, b_sb), rgb[b], b_w); Q = d.inference(n_iterations); MAP = argmax(Q)
 * (examples/inference.py:69-122; Potts compatibilities, NORMALIZE_SYMMETRIC) -- with identical results, but chunks of
 * about 4 MPix of images are run as ONE model ("super-lattice"): the images are stacked, both lattices are built once per
 * chunk with the image index as an extra key coordinate, and every kernel runs once over the chunk.  n_workers host
 * threads (0 = 3 per GPU) each drive their own chunk: one chunk at a time owns a GPU's SMs, the uploads and the
 * host-side delivery of the others run beside it.
 * ($DCRF_B200_BATCH_MODE=0, or W*H not a multiple of 4: one model per image from 4 workers per GPU, as in round 1.)
 * labels[b] receives N int16; q_out (may be NULL) points to B arrays of L*N floats that receive Q in the (L, N) row-major
 * layout of np.array(Q).
 * Several GPUs (replicas, no collective): devices[0..n_devices) lists them and worker k drives devices[k %
 * n_devices]; with devices == NULL everything runs on `device` (-1 = default). */
dcrf_status dcrf_batch_map_2d(int W, int H, int L, int device, int B, const float* const* unary_LxN,
                              const uint8_t* const* rgb_HxWx3, float g_sx, float g_sy, float g_w, float b_sx, float b_sy,
                              float b_sr, float b_sg, float b_sb, float b_w, int n_iterations, int n_workers,
                              int16_t* const* labels, float* const* q_out, const int* devices, int n_devices);

/* ---- matrices: the eigen.MatrixXf / eigen.VectorXf objects of the reference -------------------- */
/* replaces eigen.MatrixXf() (empty 0x0)               pydensecrf/eigen.pyx:62-66 */
dcrf_status dcrf_mat_create_empty(dcrf_mat_t** out);
/* replaces c_buf2matf(&arr[0,0], h, w)                pydensecrf/eigen.pyx:52 (eigen_impl.cpp:19-25): row-major in,
 * stored pixel-major (col-major).  Works without a GPU. */
dcrf_status dcrf_mat_from_host(const float* rowmajor, int rows, int cols, dcrf_mat_t** out);
/* replaces c_buf2vecf(&arr[0], n)                     pydensecrf/eigen.pyx:7 (eigen_impl.cpp:9-12): n x 1 */
dcrf_status dcrf_vec_from_host(const float* data, int n, dcrf_mat_t** out);
void dcrf_mat_destroy(dcrf_mat_t* a);
int dcrf_mat_rows(const dcrf_mat_t* a);
int dcrf_mat_cols(const dcrf_mat_t* a);
/* 1 while the matrix' current contents live in device memory (results of inference / stepInference that no host view
 * has claimed), 0 otherwise.  Never allocates. */
int dcrf_mat_on_device(const dcrf_mat_t* a);
/* replaces self.m.data() in __getbuffer__             pydensecrf/eigen.pyx:85: pointer to rows*cols floats, col-major
 * (strides (4, 4*rows)), valid until the matrix is destroyed.  Downloads from the device first if the
 * device copy is newer.  The caller may write through it: call dcrf_mat_release_host() when the view
 * ends so the next device use re-uploads. */
dcrf_status dcrf_mat_host_ptr(dcrf_mat_t* a, float** ptr);
/* Column-wise argmax of a matrix whose current contents are on the device (a result of inference / stepInference):
 * labels_cols receives `cols` labels of label_bytes bytes each (2: int16, 8: int64 = what np.argmax returns), first
 * maximum wins -- what np.argmax(Q, axis=0) computes (examples/inference.py:122, README.md:199), without moving Q to the
 * host.  DCRF_EINVAL if the matrix only exists on the host. */
dcrf_status dcrf_mat_argmax(dcrf_mat_t* a, void* labels_cols, int label_bytes);
/* device view of a result matrix for GPU-resident consumers (exposed as __cuda_array_interface__ by the binding):
 * cols rows of *pitch_floats floats, the first `rows` of each row are the labels of one variable.  Valid until the
 * matrix is destroyed or written again; fails if the current contents only exist on the host. */
dcrf_status dcrf_mat_device_ptr(dcrf_mat_t* a, float** ptr, int* pitch_floats, int* device);
/* a buffer view ended (the .pyx's __releasebuffer__, pydensecrf/eigen.pyx:97): the host copy may have been written */
void dcrf_mat_release_host(dcrf_mat_t* a);

/* ---- page-locked host memory ------------------------------------------------------------------- */
/* Host arrays passed to this library may live anywhere; arrays in page-locked memory are transferred by
 * DMA at full PCIe rate (pageable ones are staged by the driver at a fraction of it).  These two hand out
 * and take back page-locked blocks from the library's pool (pydensecrf_b200.utils.pinned_empty wraps them). */
dcrf_status dcrf_host_alloc(size_t bytes, void** out);
/* The library keeps freed device buffers (result matrices, batch slabs; up to 48 GB), page-locked host buffers (up to
 * 8 GB) and the stream-ordered pool's memory for reuse.  This returns all of it that is not in use to the driver
 * (the same happens automatically when a device allocation fails). */
void dcrf_empty_cache(void);
void dcrf_host_free(void* p, size_t bytes);
/* Page-locked memory for a RESULT (label map, copy of Q) the caller may keep: a pooled block when one is free or fewer than
 * eight blocks of this size are out, else *out = NULL and the caller uses ordinary memory (pinning costs ~0.5 ms per MB; a
 * caller who keeps every result would pay it per call).  Released with dcrf_host_free. */
dcrf_status dcrf_host_alloc_result(size_t bytes, void** out);

/* ---- measurement (bench.py) ------------------------------------------------------------------- */
dcrf_status dcrf_synchronize(dcrf_t* m);
/* CUDA events on the model's own stream: start, then stop returns elapsed milliseconds */
dcrf_status dcrf_timer_start(dcrf_t* m);
dcrf_status dcrf_timer_stop(dcrf_t* m, float* ms);
/* per-kernel-class timing: when enabled every launch is bracketed by events on the model's stream.
 * dcrf_profile_read fills up to `cap` entries; returns the count via *n. */
typedef struct {
  char name[32];
  int launches;
  float ms_total;
} dcrf_prof_entry;
dcrf_status dcrf_profile_enable(dcrf_t* m, int on);
dcrf_status dcrf_profile_read(dcrf_t* m, dcrf_prof_entry* out, int cap, int* n, int reset);
/* kernels launched by this model since creation (bench.py's gpu_launches) */
int64_t dcrf_launch_count(const dcrf_t* m);
/* write a >L2-sized scratch buffer on the model's stream (timing hygiene between iterations) */
dcrf_status dcrf_flush_l2(dcrf_t* m);
/* Self-test of the restated float arithmetic on the device (test infrastructure): which = 0 compares the packed
 * two-element restatement of Eigen's pexp with the scalar one for every float in [-100, 0] (n ignored); which = 1
 * compares the shared-reciprocal column division with IEEE division on n random operand pairs.  *mismatches
 * receives the number of results whose bits differ (must be 0). */
dcrf_status dcrf_debug_selftest(dcrf_t* m, int which, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* upload a host unary into a device-resident matrix / run inference entirely on device buffers:
 * same as dcrf_inference but reuses *Q if it already has the right shape (no allocation in the timed loop) */
dcrf_status dcrf_inference_into(dcrf_t* m, int n_iterations, dcrf_mat_t* Q);

/* ---- introspection for the parity tests -------------------------------------------------------- */
int dcrf_num_pairwise(const dcrf_t* m);
/* lattice of pairwise term k: feature dims d, pixels N, vertices V (incl. phantom vertices) */
dcrf_status dcrf_lattice_info(const dcrf_t* m, int k, int* d, int* N, int* V);
/* any pointer may be NULL.  offset/bary: N*(d+1) (vertex id / barycentric weight per pixel corner, ids in the
 * reference's first-touch order); keys: V*d int16; nbr: (d+1)*V*2 (-1 = absent); norm: N. */
dcrf_status dcrf_lattice_dump(dcrf_t* m, int k, int32_t* offset, float* bary, int16_t* keys, int32_t* nbr,
                              float* norm);
/* Permutohedral::compute of term k (permutohedral.cpp:596-604) on a host array, NxL pixel-major in/out;
 * no normalisation, no compatibility.  reverse != 0 blurs the axes in descending order (the `reverse` argument of
 * the reference, used by applyTranspose). */
dcrf_status dcrf_lattice_filter(dcrf_t* m, int k, const float* in_NxL, int L, int reverse, float* out_NxL);
/* PairwisePotential::apply of term k (pairwise.cpp:173-178) on a device matrix -> new matrix */
dcrf_status dcrf_pairwise_apply(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out);
/* ---- learning-path building blocks (not wrapped by densecrf.pyx, densecrf.pxd:61-76) ----------------------
 * PairwisePotential::applyTranspose (pairwise.cpp:179-183): filter with the normalisations swapped and the blur axes in
 * reverse order, then the (symmetric) compatibility. */
dcrf_status dcrf_pairwise_apply_transpose(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out);
/* Permutohedral::gradient (permutohedral.cpp:612-695) on term k's lattice: df[i*d + j] = d(a^T K b)/d feature j of
 * variable i, for pixel-major a, b (N rows of L floats, the layout of MatrixXf::data()); df: N rows of d floats.
 * DenseKernel::kernelGradient (pairwise.cpp:82-86) is exactly this call. */
dcrf_status dcrf_lattice_gradient(dcrf_t* m, int k, const float* a_NxL, const float* b_NxL, int L, float* df_Nxd);
/* DenseCRF::unaryEnergy / pairwiseEnergy (densecrf.cpp:141-177, densecrf.h:88-91; not wrapped by densecrf.pyx): the energy
 * every variable contributes under a labelling -- unary(l_i, i), and -0.5 * (apply of the one-hot belief)(l_i, i) for
 * pairwise term `term` (-1: summed over all terms in order).  Labels outside [0, nlabels) give 0.  out_N: N floats. */
dcrf_status dcrf_unary_energy(dcrf_t* m, const int16_t* labels_N, float* out_N);
dcrf_status dcrf_pairwise_energy(dcrf_t* m, const int16_t* labels_N, int term, float* out_N);
/* DenseCRF::gradient (densecrf.cpp:237-297, declared densecrf.h:106): runs n_iterations of mean field keeping every Q,
 * evaluates the objective on the last one and back-propagates through the iterations.  Any gradient pointer may be NULL
 * (that gradient is not computed); each receives as many floats as dcrf_num_parameters reports for its group, in the
 * reference's packing.  *value receives ObjectiveFunction::evaluate's return value.  Sums over the variables are taken in
 * double in a fixed order (deterministic); against the reference's float Eigen reductions the gradients agree to ~1e-4
 * relative (tests/test_gradient.py states the tolerances). */
typedef enum { DCRF_OBJECTIVE_LOGLIKELIHOOD = 0, DCRF_OBJECTIVE_HAMMING = 1, DCRF_OBJECTIVE_IOU = 2 } dcrf_objective_kind;
typedef struct {
  int kind;                  /* dcrf_objective_kind: LogLikelihood / Hamming / IntersectionOverUnion (objective.h:38-74) */
  const int16_t* gt;         /* ground truth, one label per variable (VectorXs); labels outside [0, nlabels) are ignored */
  float robust;              /* LogLikelihood(gt, robust) (objective.cpp:35) */
  const float* class_weight; /* Hamming(gt, w) (objective.cpp:65): n_class_weight floats, must cover every label in gt */
  int n_class_weight;
} dcrf_objective;
dcrf_status dcrf_gradient(dcrf_t* m, int n_iterations, const dcrf_objective* objective, double* value, float* unary_grad,
                          float* lbl_cmp_grad, float* kernel_grad);
/* The three parameter vectors of a model (densecrf.h:109-124; densecrf.cpp:298-375):
 *   DCRF_PARAMS_UNARY               unaryParameters / setUnaryParameters: the logistic unary's L matrix, column-major
 *                                   (unary.cpp:53-64); empty for a constant unary
 *   DCRF_PARAMS_LABEL_COMPATIBILITY labelCompatibilityParameters / set...: per term Potts [w], Diagonal w[L], Matrix the
 *                                   upper triangle row by row (labelcompatibility.cpp:49-99), concatenated over the terms
 *   DCRF_PARAMS_KERNEL              kernelParameters / set...: per term nothing (CONST_KERNEL), the d feature scales
 *                                   (DIAG_KERNEL) or the d x d feature transform, column-major (FULL_KERNEL)
 *                                   (pairwise.cpp:129-151); setting them rebuilds the term's lattice from the transformed
 *                                   features, as DenseKernel::setParameters does */
typedef enum { DCRF_PARAMS_UNARY = 0, DCRF_PARAMS_LABEL_COMPATIBILITY = 1, DCRF_PARAMS_KERNEL = 2 } dcrf_param_group;
dcrf_status dcrf_num_parameters(dcrf_t* m, int group, int* n);
dcrf_status dcrf_get_parameters(dcrf_t* m, int group, float* out);
dcrf_status dcrf_set_parameters(dcrf_t* m, int group, const float* v, int n);
/* 0 = pick automatically, 1 = force the staged (unfused) iteration, 2 = force the fused iteration (fails if
 * the model has a term the fused kernel cannot take) */
dcrf_status dcrf_set_iteration_mode(dcrf_t* m, int mode);
/* 0 (default) = bit-faithful softmax: Eigen's packet exp polynomial, Eigen's summation order, IEEE division,
 * so Q equals the reference's to the last bit except for rare last-ulp differences in the L%4 tail labels;
 * 1 = fast softmax in the fused kernel: ex2-based exp, tree sum, reciprocal multiply (~1e-6 relative). */
dcrf_status dcrf_set_softmax_mode(dcrf_t* m, int fast);
/* klDivergence's pairwise sums.  0 (default): double accumulation through a fixed reduction tree -- deterministic, and
 * within ~1e-4 relative of the reference, whose own float reduction (densecrf.cpp:232) is the less accurate side.
 * 1: the reference's reduction order restated (eight sequential float chains, Eigen's redux): the pairwise terms come
 * out bit-exact, the whole KL within 1e-9 relative; about 10 ms per term at 1080p instead of 0.1 ms. */
dcrf_status dcrf_set_kl_mode(dcrf_t* m, int mode);
/* Which build of the iteration kernels runs (bit mask; masks 0-3 give identical results): 1 = wide-lane
 * update kernel (one 256-bit access per lane, csrc/filter_wide.cu), 2 = the update kernel hands Q * diag(norm_k) to the
 * next iteration's splats, 4 = the pre-normalisation folded into the splat's contribution weights ((w * norm) * Q instead
 * of the reference's w * (norm * Q): ~1e-5 on Q after 10 iterations, 4.5 % faster at 1080p; opt-in, overrides 2).
 * Default: $DCRF_B200_VARIANT, else 3.  For A/B timing and tests. */
dcrf_status dcrf_set_kernel_variant(dcrf_t* m, int mask);
/* MatrixCompatibility::apply (labelcompatibility.cpp:82-84), out = W*Q per pixel:
 * 0 (default) = tcgen05 tensor cores with 3xTF32 split operands from 64 labels up (~1e-6 relative), the exact-order
 *               CUDA-core kernel below that; 1 = always the exact-order kernel (Eigen's GEBP summation order, bit
 *               faithful); 2 = tensor cores whenever the shape allows (16 <= labels <= 256). */
dcrf_status dcrf_set_matrix_mode(dcrf_t* m, int mode);

#ifdef __cplusplus
}
#endif
#endif /* DCRF_B200_H */
