// TEST INFRASTRUCTURE ONLY -- never linked into or imported by the product.
//
// C shim around the UNMODIFIED reference C++ (compiled from the sources where
// they lie under /root/reference by oracle/Makefile; outputs go to oracle/_ref/).
// It exposes the reference's own Permutohedral / DenseCRF objects through the
// same flat C interface that oracle/dcrf_oracle.c (the restatement) exposes
// with the orc_ prefix, so tests can drive both with one ctypes wrapper and
// pin the restatement against the real thing.
//
// Layout conventions at this interface (identical to the orc_ interface):
//   * "LxN row-major"  : what numpy hands to pydensecrf (label planes).
//   * "NxL pixel-major": the memory of an Eigen::MatrixXf(L, N) (col-major),
//     i.e. what pydensecrf's MatrixXf buffer exposes (eigen.pyx:58-98).
#include <cstdint>
#include <cstring>
#include "densecrf.h"
#include "permutohedral.h"
#include "pairwise.h"
#include "labelcompatibility.h"
#include "unary.h"
#include "objective.h"

void expAndNormalize(MatrixXf& out, const MatrixXf& in);  // densecrf.cpp:98 (non-static free function)

namespace {
typedef Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RowMat;

MatrixXf from_rowmajor(const float* p, int rows, int cols) {
  // same conversion pydensecrf does at its boundary (eigen_impl.cpp:19-25)
  return Eigen::Map<const RowMat>(p, rows, cols);
}

// Permutohedral keeps its tables protected (permutohedral.h:43-53); a subclass can read them.
struct LatticeProbe : public Permutohedral {
  int n() const { return N_; }
  int d() const { return d_; }
  int nvert() const { return M_; }
  const std::vector<int>& off() const { return offset_; }
  const std::vector<float>& bary() const { return barycentric_; }
  const std::vector<int>& rnk() const { return rank_; }
  void nbr(int j, int v, int* n1, int* n2) const {
    *n1 = blur_neighbors_[j * M_ + v].n1;
    *n2 = blur_neighbors_[j * M_ + v].n2;
  }
};

LabelCompatibility* make_compat(int kind, const float* data, int L) {
  if (kind == 0) return new PottsCompatibility(data[0]);
  if (kind == 1) {
    VectorXf v = Eigen::Map<const VectorXf>(data, L);
    return new DiagonalCompatibility(v);
  }
  return new MatrixCompatibility(from_rowmajor(data, L, L));
}

struct Model {
  DenseCRF* crf;
  DenseCRF2D* crf2d;
  int N, L;
  MatrixXf Q, t1, t2;  // the caller-owned matrices of the startInference / stepInference protocol (densecrf.pyx:94-98)
};
}  // namespace

extern "C" {

void* ref_lattice_build(const float* feat_dxN, int d, int N) {
  LatticeProbe* p = new LatticeProbe();
  p->init(from_rowmajor(feat_dxN, d, N));
  return p;
}
void ref_lattice_free(void* h) { delete static_cast<LatticeProbe*>(h); }
int ref_lattice_nvert(void* h) { return static_cast<LatticeProbe*>(h)->nvert(); }
void ref_lattice_dump(void* h, int* offset, float* bary, int* rank, int* nbr) {
  LatticeProbe* p = static_cast<LatticeProbe*>(h);
  const int n = p->n() * (p->d() + 1);
  if (offset) std::memcpy(offset, p->off().data(), sizeof(int) * n);
  if (bary) std::memcpy(bary, p->bary().data(), sizeof(float) * n);
  if (rank) std::memcpy(rank, p->rnk().data(), sizeof(int) * n);
  if (nbr) {
    const int V = p->nvert();
    for (int j = 0; j <= p->d(); j++)
      for (int v = 0; v < V; v++) p->nbr(j, v, &nbr[2 * (j * V + v)], &nbr[2 * (j * V + v) + 1]);
  }
}
void ref_lattice_filter(void* h, float* out_NxL, const float* in_NxL, int L, int reverse) {
  LatticeProbe* p = static_cast<LatticeProbe*>(h);
  MatrixXf in = Eigen::Map<const MatrixXf>(in_NxL, L, p->n());
  MatrixXf out;
  p->compute(out, in, reverse != 0);
  std::memcpy(out_NxL, out.data(), sizeof(float) * (size_t)L * p->n());
}

void ref_lattice_gradient(void* h, float* df_Nxd, const float* a_NxL, const float* b_NxL, int L) {
  static_cast<LatticeProbe*>(h)->gradient(df_Nxd, a_NxL, b_NxL, L);
}

void ref_exp_and_normalize(float* out_NxL, const float* in_NxL, int L, int N) {
  MatrixXf in = Eigen::Map<const MatrixXf>(in_NxL, L, N);
  MatrixXf out;
  expAndNormalize(out, in);
  std::memcpy(out_NxL, out.data(), sizeof(float) * (size_t)L * N);
}

void* ref_crf_new(int N, int L) {
  Model* m = new Model();
  m->crf = new DenseCRF(N, L);
  m->crf2d = 0;
  m->N = N;
  m->L = L;
  return m;
}
void* ref_crf_new_2d(int W, int H, int L) {
  Model* m = new Model();
  m->crf2d = new DenseCRF2D(W, H, L);
  m->crf = m->crf2d;
  m->N = W * H;
  m->L = L;
  return m;
}
void ref_crf_free(void* h) {
  Model* m = static_cast<Model*>(h);
  delete m->crf;
  delete m;
}
void ref_crf_set_unary(void* h, const float* u_LxN) {
  Model* m = static_cast<Model*>(h);
  m->crf->setUnaryEnergy(from_rowmajor(u_LxN, m->L, m->N));
}
void ref_crf_set_unary_logistic(void* h, const float* Lm_LxF, int F, const float* f_FxN) {
  Model* m = static_cast<Model*>(h);
  m->crf->setUnaryEnergy(from_rowmajor(Lm_LxF, m->L, F), from_rowmajor(f_FxN, F, m->N));
}
void ref_crf_add_pairwise(void* h, const float* feat_dxN, int d, int compat_kind, const float* compat,
                          int ktype, int ntype) {
  Model* m = static_cast<Model*>(h);
  m->crf->addPairwiseEnergy(from_rowmajor(feat_dxN, d, m->N), make_compat(compat_kind, compat, m->L),
                            (KernelType)ktype, (NormalizationType)ntype);
}
int ref_crf_add_gaussian(void* h, float sx, float sy, int compat_kind, const float* compat, int ktype,
                         int ntype) {
  Model* m = static_cast<Model*>(h);
  if (!m->crf2d) return -1;
  m->crf2d->addPairwiseGaussian(sx, sy, make_compat(compat_kind, compat, m->L), (KernelType)ktype,
                                (NormalizationType)ntype);
  return 0;
}
int ref_crf_add_bilateral(void* h, float sx, float sy, float sr, float sg, float sb, const uint8_t* im,
                          int compat_kind, const float* compat, int ktype, int ntype) {
  Model* m = static_cast<Model*>(h);
  if (!m->crf2d) return -1;
  m->crf2d->addPairwiseBilateral(sx, sy, sr, sg, sb, im, make_compat(compat_kind, compat, m->L),
                                 (KernelType)ktype, (NormalizationType)ntype);
  return 0;
}
void ref_crf_inference(void* h, int n_iter, float* Q_NxL) {
  Model* m = static_cast<Model*>(h);
  MatrixXf Q = m->crf->inference(n_iter);
  std::memcpy(Q_NxL, Q.data(), sizeof(float) * (size_t)m->L * m->N);
}
void ref_crf_start(void* h, float* Q_NxL) {
  Model* m = static_cast<Model*>(h);
  MatrixXf Q = m->crf->startInference();
  std::memcpy(Q_NxL, Q.data(), sizeof(float) * (size_t)m->L * m->N);
}
void ref_crf_step(void* h, float* Q_NxL) {
  Model* m = static_cast<Model*>(h);
  MatrixXf Q = Eigen::Map<const MatrixXf>(Q_NxL, m->L, m->N), t1, t2;
  m->crf->stepInference(Q, t1, t2);
  std::memcpy(Q_NxL, Q.data(), sizeof(float) * (size_t)m->L * m->N);
}
// The reference's stepwise protocol exactly as its Python wrapper drives it: Q, tmp1, tmp2 = startInference() once, then
// stepInference(Q, tmp1, tmp2) on the same three matrices -- no copies in or out (used as the timed CPU arm).
void ref_crf_persist_start(void* h) {
  Model* m = static_cast<Model*>(h);
  m->Q = m->crf->startInference();
  m->t1 = MatrixXf();
  m->t2 = MatrixXf();
}
void ref_crf_persist_step(void* h) {
  Model* m = static_cast<Model*>(h);
  m->crf->stepInference(m->Q, m->t1, m->t2);
}
void ref_crf_persist_get(void* h, float* Q_NxL) {
  Model* m = static_cast<Model*>(h);
  std::memcpy(Q_NxL, m->Q.data(), sizeof(float) * (size_t)m->L * m->N);
}
double ref_crf_kl(void* h, const float* Q_NxL) {
  Model* m = static_cast<Model*>(h);
  MatrixXf Q = Eigen::Map<const MatrixXf>(Q_NxL, m->L, m->N);
  return m->crf->klDivergence(Q);
}
// message of pairwise term k alone: PairwisePotential::apply (pairwise.cpp:173-178) is not reachable
// through DenseCRF's public interface, so build a throw-away potential from the same arguments.
void ref_pairwise_apply(const float* feat_dxN, int d, int N, int L, int compat_kind, const float* compat,
                        int ktype, int ntype, const float* Q_NxL, float* out_NxL) {
  PairwisePotential pp(from_rowmajor(feat_dxN, d, N), make_compat(compat_kind, compat, L), (KernelType)ktype,
                       (NormalizationType)ntype);
  MatrixXf Q = Eigen::Map<const MatrixXf>(Q_NxL, L, N), out;
  pp.apply(out, Q);
  std::memcpy(out_NxL, out.data(), sizeof(float) * (size_t)L * N);
}
// DenseCRF::unaryEnergy / pairwiseEnergy (densecrf.cpp:141-177)
void ref_crf_unary_energy(void* h, const short* labels, float* out_N) {
  Model* m = static_cast<Model*>(h);
  VectorXf r = m->crf->unaryEnergy(VectorXs(Eigen::Map<const VectorXs>(labels, m->N)));
  std::memcpy(out_N, r.data(), sizeof(float) * m->N);
}
void ref_crf_pairwise_energy(void* h, const short* labels, int term, float* out_N) {
  Model* m = static_cast<Model*>(h);
  VectorXf r = m->crf->pairwiseEnergy(VectorXs(Eigen::Map<const VectorXs>(labels, m->N)), term);
  std::memcpy(out_N, r.data(), sizeof(float) * m->N);
}
// ---- learning path: the reference's own DenseCRF::gradient and parameter vectors (densecrf.h:106-124) ----------------
// group: 0 unary, 1 label compatibility, 2 kernel
static VectorXf ref_params(Model* m, int group) {
  return group == 0 ? m->crf->unaryParameters() : group == 1 ? m->crf->labelCompatibilityParameters() : m->crf->kernelParameters();
}
int ref_crf_num_params(void* h, int group) { return (int)ref_params(static_cast<Model*>(h), group).rows(); }
void ref_crf_get_params(void* h, int group, float* out) {
  VectorXf v = ref_params(static_cast<Model*>(h), group);
  if (v.rows()) std::memcpy(out, v.data(), sizeof(float) * v.rows());
}
void ref_crf_set_params(void* h, int group, const float* v, int n) {
  Model* m = static_cast<Model*>(h);
  VectorXf p = Eigen::Map<const VectorXf>(v, n);
  if (group == 0) m->crf->setUnaryParameters(p);
  else if (group == 1) m->crf->setLabelCompatibilityParameters(p);
  else m->crf->setKernelParameters(p);
}
// obj_kind: 0 LogLikelihood(gt, robust), 1 Hamming(gt, weights), 2 IntersectionOverUnion(gt); a null gradient pointer
// skips that gradient, as a null VectorXf* does in the reference
double ref_crf_gradient(void* h, int n_iter, int obj_kind, const short* gt, float robust, const float* cw, int ncw,
                        float* unary_grad, float* lbl_cmp_grad, float* kernel_grad) {
  Model* m = static_cast<Model*>(h);
  VectorXs g = Eigen::Map<const VectorXs>(gt, m->N);
  ObjectiveFunction* obj;
  if (obj_kind == 0) obj = new LogLikelihood(g, robust);
  else if (obj_kind == 1) obj = new Hamming(g, VectorXf(Eigen::Map<const VectorXf>(cw, ncw)));
  else obj = new IntersectionOverUnion(g);
  VectorXf ug, lg, kg;
  double r = m->crf->gradient(n_iter, *obj, unary_grad ? &ug : 0, lbl_cmp_grad ? &lg : 0, kernel_grad ? &kg : 0);
  delete obj;
  if (unary_grad && ug.rows()) std::memcpy(unary_grad, ug.data(), sizeof(float) * ug.rows());
  if (lbl_cmp_grad && lg.rows()) std::memcpy(lbl_cmp_grad, lg.data(), sizeof(float) * lg.rows());
  if (kernel_grad && kg.rows()) std::memcpy(kernel_grad, kg.data(), sizeof(float) * kg.rows());
  return r;
}
// Hamming(gt, class_weight_pow)'s class weights (objective.cpp:51-63): the class keeps them private, so they are read
// back through evaluate() on a Q of ones (d_mul_Q(gt_i, i) = class_weight[gt_i])
void ref_hamming_weights(const short* gt, int N, float pw, int L, float* out_L) {
  VectorXs g = Eigen::Map<const VectorXs>(gt, N);
  Hamming hm(g, pw);
  MatrixXf Q = MatrixXf::Ones(L, N), d;
  hm.evaluate(d, Q);
  for (int l = 0; l < L; l++) out_L[l] = 0.f;
  for (int i = 0; i < N; i++)
    if (gt[i] >= 0 && gt[i] < L) out_L[gt[i]] = d(gt[i], i);
}
const char* ref_describe() { return "reference C++ compiled from /root/reference (g++ -O2 -DNDEBUG, SSE2)"; }
}
