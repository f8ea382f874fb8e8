/*
 * dcrf_oracle.c -- CPU restatement of the reference's mean-field hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pydensecrf_b200/ may import, link or execute this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the CPU figure.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here bit-for-bit
 * (lattice tables, filter output, softmax, full inference Q) against the UNMODIFIED reference
 * C++ compiled by oracle/Makefile into oracle/_ref/libref.so, and tests/golden/ holds vectors
 * generated from that reference (tests/golden/make_golden.py) for boxes where _ref is absent.
 *
 * This is a plain-C restatement of the algorithm, written from the behaviour of the reference,
 * not a copy: scalar code that reproduces, operation for operation, the float arithmetic the
 * x86-64 build of the reference performs (SSE2 path, no FMA, round-to-nearest-even; SURVEY.md
 * section 8a quirk list).  All file:line citations are relative to
 * /root/reference/pydensecrf/densecrf/ .
 *
 * Layouts at this interface:
 *   "LxN row-major"   what numpy hands to pydensecrf (one plane per label)
 *   "NxL pixel-major" the memory of an Eigen::MatrixXf(L,N): the L labels of a pixel contiguous
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Vertex table: key (d int16 lattice coordinates) -> dense vertex id in first-touch order.    */
/* Restates HashTable (src/permutohedral.cpp:54-131).  Only two properties are observable by   */
/* the rest of the algorithm: ids are handed out in insertion order (:110) and lookups of      */
/* absent keys give -1 (:113); the hash function and growth policy are free, so they are ours. */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int d;
  size_t cap, count; /* cap is a power of two */
  int *slot;         /* cap entries, -1 = empty */
  short *keys;       /* count*d, in id order */
  size_t keys_cap;
} vtable;

static uint64_t vt_hash(const short *k, int d) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < d; i++) {
    h ^= (uint16_t)k[i];
    h *= 1099511628211ull;
  }
  return h ^ (h >> 29);
}
static void vt_init(vtable *t, int d, size_t expect) {
  t->d = d;
  t->cap = 64;
  while (t->cap < 2 * expect) t->cap <<= 1;
  t->count = 0;
  t->slot = (int *)malloc(t->cap * sizeof(int));
  memset(t->slot, 0xff, t->cap * sizeof(int));
  t->keys_cap = expect > 16 ? expect : 16;
  t->keys = (short *)malloc(t->keys_cap * (size_t)(d > 0 ? d : 1) * sizeof(short));
}
static void vt_free(vtable *t) {
  free(t->slot);
  free(t->keys);
}
static void vt_rehash(vtable *t) {
  free(t->slot);
  t->cap <<= 1;
  t->slot = (int *)malloc(t->cap * sizeof(int));
  memset(t->slot, 0xff, t->cap * sizeof(int));
  for (size_t e = 0; e < t->count; e++) {
    size_t h = vt_hash(t->keys + e * t->d, t->d) & (t->cap - 1);
    while (t->slot[h] >= 0) h = (h + 1) & (t->cap - 1);
    t->slot[h] = (int)e;
  }
}
static int vt_find(vtable *t, const short *k, int create) {
  size_t h = vt_hash(k, t->d) & (t->cap - 1);
  for (;;) {
    int e = t->slot[h];
    if (e < 0) {
      if (!create) return -1;
      if (t->count == t->keys_cap) {
        t->keys_cap *= 2;
        t->keys = (short *)realloc(t->keys, t->keys_cap * (size_t)t->d * sizeof(short));
      }
      memcpy(t->keys + t->count * t->d, k, (size_t)t->d * sizeof(short));
      t->slot[h] = (int)t->count;
      int id = (int)t->count++;
      if (2 * t->count >= t->cap) vt_rehash(t);
      return id;
    }
    if (memcmp(t->keys + (size_t)e * t->d, k, (size_t)t->d * sizeof(short)) == 0) return e;
    h = (h + 1) & (t->cap - 1);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Lattice                                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int N, d, V;
  int *offset; /* [Npad*(d+1)] vertex id of (pixel, corner); Npad = N rounded up to 4 */
  float *bary; /* [Npad*(d+1)] barycentric weight */
  int *rank;   /* [Npad*(d+1)] (only the learning path reads it; dumped for completeness) */
  int *nbr;    /* [(d+1)][V][2]: neighbour ids -/+ along each axis, -1 = not in the lattice */
  short *keys; /* [V][d] */
} orc_lattice;

#define ORC_MAXD 64

/*
 * One pixel of Permutohedral::init, SSE flavour (src/permutohedral.cpp:191-277), one lane at a
 * time.  Every operation is a separately rounded float op, in the order the intrinsics run.
 *   f[d]      feature vector
 *   scale[d]  (float)( 1.0/sqrt((i+2)(i+1)) * (double)(float)(sqrt(2.0/3.0)*(d+1)) )   (:179-182)
 * Outputs rem0[d+1] (coordinates of the enclosing simplex' zero-colour vertex), rnk[d+1]
 * (position of each residual in sorted order) and w[d+1] barycentric weights.
 */
static void orc_locate(const float *f, int d, const float *scale, float *rem0, float *rnk, float *w) {
  float el[ORC_MAXD + 1], b[ORC_MAXD + 2];
  const float inv_dp1 = 1.0f / (float)(d + 1); /* :148 */
  const float dp1 = (float)(d + 1);

  /* elevate onto the hyperplane sum(x)=0  (:201-207): cf = f*scale; el[j] = sm - j*cf; sm += cf */
  float sm = 0.0f;
  for (int j = d; j > 0; j--) {
    float cf = f[j - 1] * scale[j - 1];
    float jc = (float)j * cf;
    el[j] = sm - jc;
    sm = sm + cf;
  }
  el[0] = sm;

  /* nearest multiple of (d+1) per coordinate, ties to even (cvtps_epi32 under MXCSR RN; :210-220) */
  float sum = 0.0f;
  for (int i = 0; i <= d; i++) {
    float v = inv_dp1 * el[i];
    v = (float)(int32_t)lrintf(v);
    rem0[i] = v * dp1;
    sum = sum + v;
  }

  /* rank the residuals; on a tie the later coordinate ranks higher (strict <, :223-233) */
  for (int i = 0; i <= d; i++) rnk[i] = 0.0f;
  for (int i = 0; i < d; i++) {
    float di = el[i] - rem0[i];
    for (int j = i + 1; j <= d; j++) {
      float dj = el[j] - rem0[j];
      float c = (di < dj) ? 1.0f : 0.0f;
      rnk[i] = rnk[i] + c;
      rnk[j] = rnk[j] + (1.0f - c);
    }
  }

  /* rounded point off the plane (sum != 0): shift ranks, wrapping by +-(d+1)  (:236-242) */
  for (int i = 0; i <= d; i++) {
    rnk[i] = rnk[i] + sum;
    float add = (rnk[i] < 0.0f) ? dp1 : 0.0f;
    float sub = (rnk[i] >= dp1) ? dp1 : 0.0f;
    float adj = add - sub;
    rnk[i] = rnk[i] + adj;
    rem0[i] = rem0[i] + adj;
  }

  /* barycentric weights, accumulated in coordinate order  (:245-263) */
  for (int i = 0; i < d + 2; i++) b[i] = 0.0f;
  for (int i = 0; i <= d; i++) {
    float v = (el[i] - rem0[i]) * inv_dp1;
    int p = (int)((float)d - rnk[i]);
    b[p] = b[p] + v;
    b[p + 1] = b[p + 1] - v;
  }
  b[0] = b[0] + (1.0f + b[d + 1]);
  for (int i = 0; i <= d; i++) w[i] = b[i];
}

/* canonical simplex entry: coordinate with rank c of the vertex with remainder r  (:171-176) */
static inline int orc_canon(int d, int r, int c) { return (c <= d - r) ? r : r - (d + 1); }

void *orc_lattice_build(const float *feat_dxN, int d, int N) {
  if (d < 1 || d > ORC_MAXD || N < 0) return NULL;
  orc_lattice *lt = (orc_lattice *)calloc(1, sizeof(orc_lattice));
  const int Npad = (N + 3) & ~3; /* the reference walks pixels in blocks of 4 (:191) ... */
  const int dp = d + 1;
  lt->N = N;
  lt->d = d;
  lt->offset = (int *)calloc((size_t)(Npad + 16) * dp, sizeof(int));
  lt->bary = (float *)calloc((size_t)(Npad + 16) * dp, sizeof(float));
  lt->rank = (int *)calloc((size_t)(Npad + 16) * dp, sizeof(int));

  float scale[ORC_MAXD];
  const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)(d + 1));
  for (int i = 0; i < d; i++) scale[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);

  vtable vt;
  vt_init(&vt, d, (size_t)(N > 16 ? N : 16));
  float f[ORC_MAXD], rem0[ORC_MAXD + 1], rnk[ORC_MAXD + 1], w[ORC_MAXD + 1];
  short key[ORC_MAXD];
  for (int i = 0; i < Npad; i++) {
    /* ... and lanes past N are fed zero features yet still inserted (:196, :261-276): the
       origin simplex becomes part of the lattice whenever N%4 != 0 ("phantom vertices"). */
    for (int j = 0; j < d; j++) f[j] = (i < N) ? feat_dxN[(size_t)j * N + i] : 0.0f;
    orc_locate(f, d, scale, rem0, rnk, w);
    for (int r = 0; r <= d; r++) {
      for (int k = 0; k < d; k++) {
        float c = rem0[k] + (float)orc_canon(d, r, (int)rnk[k]);
        key[k] = (short)(int)c; /* short truncation (:270) */
      }
      lt->offset[(size_t)i * dp + r] = vt_find(&vt, key, 1);
      lt->rank[(size_t)i * dp + r] = (int)rnk[r];
      lt->bary[(size_t)i * dp + r] = w[r];
    }
  }
  lt->V = (int)vt.count;

  /* neighbours along each of the d+1 lattice axes (:305-318): key -/+ (1,..,1) with the axis
     coordinate moved by +/- d instead; the implied (d+1)-th coordinate is not part of the key */
  lt->nbr = (int *)malloc(sizeof(int) * 2 * (size_t)dp * (lt->V > 0 ? lt->V : 1));
  short n1[ORC_MAXD + 1], n2[ORC_MAXD + 1];
  for (int j = 0; j <= d; j++)
    for (int v = 0; v < lt->V; v++) {
      const short *k0 = vt.keys + (size_t)v * d;
      for (int k = 0; k < d; k++) {
        n1[k] = (short)(k0[k] - 1);
        n2[k] = (short)(k0[k] + 1);
      }
      if (j < d) {
        n1[j] = (short)(k0[j] + d);
        n2[j] = (short)(k0[j] - d);
      }
      lt->nbr[2 * ((size_t)j * lt->V + v)] = vt_find(&vt, n1, 0);
      lt->nbr[2 * ((size_t)j * lt->V + v) + 1] = vt_find(&vt, n2, 0);
    }
  lt->keys = (short *)malloc(sizeof(short) * (size_t)d * (lt->V > 0 ? lt->V : 1));
  memcpy(lt->keys, vt.keys, sizeof(short) * (size_t)d * lt->V);
  vt_free(&vt);
  return lt;
}

void orc_lattice_free(void *h) {
  orc_lattice *lt = (orc_lattice *)h;
  if (!lt) return;
  free(lt->offset);
  free(lt->bary);
  free(lt->rank);
  free(lt->nbr);
  free(lt->keys);
  free(lt);
}
int orc_lattice_nvert(void *h) { return ((orc_lattice *)h)->V; }
void orc_lattice_dump(void *h, int *offset, float *bary, int *rank, int *nbr) {
  orc_lattice *lt = (orc_lattice *)h;
  size_t n = (size_t)lt->N * (lt->d + 1);
  if (offset) memcpy(offset, lt->offset, n * sizeof(int));
  if (bary) memcpy(bary, lt->bary, n * sizeof(float));
  if (rank) memcpy(rank, lt->rank, n * sizeof(int));
  if (nbr) memcpy(nbr, lt->nbr, sizeof(int) * 2 * (size_t)(lt->d + 1) * lt->V);
}
void orc_lattice_keys(void *h, short *keys_Vxd) {
  orc_lattice *lt = (orc_lattice *)h;
  memcpy(keys_Vxd, lt->keys, sizeof(short) * (size_t)lt->d * lt->V);
}

/*
 * Permutohedral::compute (src/permutohedral.cpp:476-610): splat, d+1 Jacobi blurs, slice.
 * The reference has two arithmetic flavours (quirk 6): rows<=2 -> seqCompute, else sseCompute.
 *   seq: blur is old + 0.5*(n1+n2) with a double literal (:505), slice multiplies (w*value)*alpha (:520)
 *   sse: blur is old + half*(n1+n2) in float (:566),             slice multiplies (w*alpha)*value (:579)
 * Row 0 of the value buffers is the all-zero stand-in for a missing neighbour.
 */
void orc_lattice_filter(void *h, float *out_NxL, const float *in_NxL, int L, int reverse) {
  orc_lattice *lt = (orc_lattice *)h;
  const int N = lt->N, d = lt->d, V = lt->V, dp = d + 1;
  const int seq = (L <= 2); /* dispatch at :596-604 */
  const size_t rows = (size_t)V + 2;
  float *val = (float *)calloc(rows * L, sizeof(float));
  float *nxt = (float *)calloc(rows * L, sizeof(float));

  for (int i = 0; i < N; i++)
    for (int j = 0; j <= d; j++) {
      const size_t o = (size_t)lt->offset[(size_t)i * dp + j] + 1;
      const float w = lt->bary[(size_t)i * dp + j];
      for (int k = 0; k < L; k++) {
        float p = w * in_NxL[(size_t)i * L + k];
        val[o * L + k] = val[o * L + k] + p;
      }
    }

  for (int jj = 0; jj <= d; jj++) {
    const int j = reverse ? d - jj : jj;
    for (int v = 0; v < V; v++) {
      const float *a = val + ((size_t)lt->nbr[2 * ((size_t)j * V + v)] + 1) * L;
      const float *b = val + ((size_t)lt->nbr[2 * ((size_t)j * V + v) + 1] + 1) * L;
      const float *o = val + ((size_t)v + 1) * L;
      float *n = nxt + ((size_t)v + 1) * L;
      if (seq)
        for (int k = 0; k < L; k++) n[k] = (float)((double)o[k] + 0.5 * (double)(a[k] + b[k]));
      else
        for (int k = 0; k < L; k++) {
          float s = a[k] + b[k];
          float hs = 0.5f * s;
          n[k] = o[k] + hs;
        }
    }
    float *t = val;
    val = nxt;
    nxt = t;
  }

  const float alpha = 1.0f / (1.0f + powf(2.0f, (float)-d)); /* :509, :571 */
  for (int i = 0; i < N; i++) {
    float *o = out_NxL + (size_t)i * L;
    /* `in` may alias `out` (pairwise.cpp:74): splat is finished, so this is safe */
    float acc[4096];
    float *ac = (L <= 4096) ? acc : (float *)malloc(sizeof(float) * L);
    for (int k = 0; k < L; k++) ac[k] = 0.0f;
    for (int j = 0; j <= d; j++) {
      const size_t r = (size_t)lt->offset[(size_t)i * dp + j] + 1;
      const float w = lt->bary[(size_t)i * dp + j];
      if (seq)
        for (int k = 0; k < L; k++) {
          float p = w * val[r * L + k];
          p = p * alpha;
          ac[k] = ac[k] + p;
        }
      else {
        const float wa = w * alpha;
        for (int k = 0; k < L; k++) {
          float p = wa * val[r * L + k];
          ac[k] = ac[k] + p;
        }
      }
    }
    for (int k = 0; k < L; k++) o[k] = ac[k];
    if (ac != acc) free(ac);
  }
  free(val);
  free(nxt);
}

/* ------------------------------------------------------------------------------------------ */
/* expAndNormalize (src/densecrf.cpp:98-106) with Eigen 3.3.9's SSE2 evaluation (quirk 9)      */
/* ------------------------------------------------------------------------------------------ */

/* One lane of Eigen's pexp<Packet4f> (include/Eigen/src/Core/arch/SSE/MathFunctions.h:107-170):
   Cephes range reduction + degree-5 polynomial, no FMA, floor via truncate-and-fix (no SSE4.1). */
static float orc_pexp_lane(float x0) {
  union {
    int32_t i;
    float f;
  } pw;
  float x = x0;
  x = (x < 88.3762626647950f) ? x : 88.3762626647950f;   /* pmin */
  x = (x > -88.3762626647949f) ? x : -88.3762626647949f; /* pmax */
  float fx = x * 1.44269504088896341f;
  fx = fx + 0.5f;
  float t = (float)(int32_t)fx; /* cvttps: toward zero */
  if (t > fx) t = t - 1.0f;
  fx = t;
  float tmp = fx * 0.693359375f;
  float z = fx * -2.12194440e-4f;
  x = x - tmp;
  x = x - z;
  z = x * x;
  float y = 1.9875691500E-4f;
  y = y * x;
  y = y + 1.3981999507E-3f;
  y = y * x;
  y = y + 8.3334519073E-3f;
  y = y * x;
  y = y + 4.1665795894E-2f;
  y = y * x;
  y = y + 1.6666665459E-1f;
  y = y * x;
  y = y + 5.0000001201E-1f;
  y = y * z;
  y = y + x;
  y = y + 1.0f;
  pw.i = ((int32_t)fx + 0x7f) << 23;
  float r = y * pw.f;
  return (r > x0) ? r : x0; /* pmax(result, _x) */
}

/*
 * Permutohedral::gradient (src/permutohedral.cpp:612-695): d/d(feature) of a^T K b.  For dir in {0,1}: splat a (dir 0) or
 * b (dir 1); blur the axes ascending (dir 0) or descending (dir 1) with the scalar loop's arithmetic (0.5 is a double
 * literal, :657); per pixel rotate the blurred corner values by the stored ranks (:660-672), multiply by the elevation
 * matrix and dot with the other operand (:674-688).  df: N x d, pixel-major; a, b: N x vs.
 */
void orc_lattice_gradient(void *h, float *df, const float *a, const float *b, int vs) {
  orc_lattice *lt = (orc_lattice *)h;
  const int N = lt->N, d = lt->d, V = lt->V, dp = d + 1;
  const size_t rows = (size_t)V + 2;
  float *val = (float *)malloc(rows * vs * sizeof(float));
  float *nxt = (float *)malloc(rows * vs * sizeof(float));
  float *r_a = (float *)malloc(sizeof(float) * (size_t)dp * vs);
  float *sm = (float *)malloc(sizeof(float) * (size_t)vs);
  float scale[ORC_MAXD];
  for (size_t e = 0; e < (size_t)N * d; e++) df[e] = 0.f;
  const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (d + 1)); /* :623 */
  for (int i = 0; i < d; i++) scale[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev); /* :625 */
  const float alpha = 1.0f / (1 + powf(2, -d)) / (d + 1); /* :629 */

  for (int dir = 0; dir < 2; dir++) {
    const float *in = dir ? b : a, *other = dir ? a : b;
    for (size_t e = 0; e < rows * vs; e++) val[e] = nxt[e] = 0;
    for (int i = 0; i < N; i++)
      for (int j = 0; j <= d; j++) {
        const size_t o = (size_t)lt->offset[(size_t)i * dp + j] + 1;
        const float w = lt->bary[(size_t)i * dp + j];
        for (int k = 0; k < vs; k++) {
          float p = w * in[(size_t)i * vs + k];
          val[o * vs + k] = val[o * vs + k] + p;
        }
      }
    for (int j = dir ? d : 0; j <= d && j >= 0; dir ? j-- : j++) {
      for (int v = 0; v < V; v++) {
        const float *o = val + ((size_t)v + 1) * vs;
        float *n = nxt + ((size_t)v + 1) * vs;
        const float *n1 = val + ((size_t)lt->nbr[2 * ((size_t)j * V + v)] + 1) * vs;
        const float *n2 = val + ((size_t)lt->nbr[2 * ((size_t)j * V + v) + 1] + 1) * vs;
        for (int k = 0; k < vs; k++) n[k] = (float)((double)o[k] + 0.5 * (double)(n1[k] + n2[k]));
      }
      float *t = val;
      val = nxt;
      nxt = t;
    }
    for (int i = 0; i < N; i++) {
      for (int e = 0; e < dp * vs; e++) r_a[e] = 0.f;
      for (int j = 0; j <= d; j++) {
        const int r0 = d - lt->rank[(size_t)i * dp + j];
        const int r1 = r0 + 1 > d ? 0 : r0 + 1;
        const size_t o0 = (size_t)lt->offset[(size_t)i * dp + r0] + 1;
        const size_t o1 = (size_t)lt->offset[(size_t)i * dp + r1] + 1;
        for (int k = 0; k < vs; k++) {
          float p0 = alpha * val[o0 * vs + k];
          r_a[j * vs + k] = r_a[j * vs + k] + p0;
          float p1 = alpha * val[o1 * vs + k];
          r_a[j * vs + k] = r_a[j * vs + k] - p1;
        }
      }
      for (int k = 0; k < vs; k++) sm[k] = r_a[k];
      for (int j = 1; j <= d; j++) {
        float grad = 0;
        for (int k = 0; k < vs; k++) {
          float jr = (float)j * r_a[j * vs + k];
          float df0 = sm[k] - jr;
          float v = scale[j - 1] * df0;
          float p = other[(size_t)i * vs + k] * v;
          grad = grad + p;
          sm[k] = sm[k] + r_a[j * vs + k];
        }
        df[(size_t)i * d + j - 1] = df[(size_t)i * d + j - 1] + grad;
      }
    }
  }
  free(val);
  free(nxt);
  free(r_a);
  free(sm);
}

/* Eigen's linear vectorised sum over a 16-byte aligned float array
   (include/Eigen/src/Core/Redux.h:213-262 with Packet4f, predux at arch/SSE/PacketMath.h:540-551):
   two packet accumulators over 8-element strides, one optional trailing packet, horizontal add
   as (a0+a2)+(a1+a3), then the scalar tail. */
static float orc_eigen_sum(const float *b, long n) {
  const long al2 = (n / 8) * 8, al = (n / 4) * 4;
  if (al == 0) {
    float r = b[0];
    for (long i = 1; i < n; i++) r = r + b[i];
    return r;
  }
  float p0[4], p1[4];
  for (int l = 0; l < 4; l++) p0[l] = b[l];
  if (al > 4) {
    for (int l = 0; l < 4; l++) p1[l] = b[4 + l];
    for (long i = 8; i < al2; i += 8)
      for (int l = 0; l < 4; l++) {
        p0[l] = p0[l] + b[i + l];
        p1[l] = p1[l] + b[i + 4 + l];
      }
    for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
    if (al > al2)
      for (int l = 0; l < 4; l++) p0[l] = p0[l] + b[al2 + l];
  }
  float t0 = p0[0] + p0[2], t1 = p0[1] + p0[3];
  float r = t0 + t1;
  for (long i = al; i < n; i++) r = r + b[i];
  return r;
}

void orc_exp_and_normalize(float *out_NxL, const float *in_NxL, int L, int N) {
  float *b = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
  const int al = (L / 4) * 4;
  for (int i = 0; i < N; i++) {
    const float *x = in_NxL + (size_t)i * L;
    float mx = x[0];
    for (int k = 1; k < L; k++)
      if (x[k] > mx) mx = x[k];
    for (int k = 0; k < L; k++) b[k] = x[k] - mx;
    for (int k = 0; k < al; k++) b[k] = orc_pexp_lane(b[k]); /* whole packets: Eigen's polynomial */
    for (int k = al; k < L; k++) b[k] = expf(b[k]);          /* L%4 tail: libm */
    const float s = orc_eigen_sum(b, L);
    float *o = out_NxL + (size_t)i * L;
    for (int k = 0; k < L; k++) o[k] = b[k] / s;
  }
  free(b);
}

/* ------------------------------------------------------------------------------------------ */
/* Model: unary + pairwise terms + mean-field driver                                           */
/* ------------------------------------------------------------------------------------------ */
enum { ORC_POTTS = 0, ORC_DIAG = 1, ORC_MATRIX = 2 };
enum { ORC_NO_NORM = 0, ORC_NORM_BEFORE = 1, ORC_NORM_AFTER = 2, ORC_NORM_SYMMETRIC = 3 };

typedef struct {
  orc_lattice *lat;
  float *norm; /* [N] */
  int ntype;
  int ckind;
  float *cw; /* 1 | L | L*L (row-major, already symmetrised) */
} orc_term;

typedef struct {
  int N, L, W, H; /* W,H = 0 for the generic model */
  float *unary;   /* NxL pixel-major, or NULL */
  int nterm;
  orc_term *term;
  float *pQ; /* the caller-owned Q of the startInference / stepInference protocol (orc_crf_persist_*) */
} orc_crf;

static float *orc_to_pixel_major(const float *rowmajor, int rows, int cols) {
  /* the row-major -> col-major copy pydensecrf performs at its boundary (eigen_impl.cpp:19-25) */
  float *p = (float *)malloc(sizeof(float) * (size_t)rows * cols + 4);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) p[(size_t)c * rows + r] = rowmajor[(size_t)r * cols + c];
  return p;
}

void *orc_crf_new(int N, int L) {
  orc_crf *m = (orc_crf *)calloc(1, sizeof(orc_crf));
  m->N = N;
  m->L = L;
  return m;
}
void *orc_crf_new_2d(int W, int H, int L) {
  orc_crf *m = (orc_crf *)orc_crf_new(W * H, L);
  m->W = W;
  m->H = H;
  return m;
}
void orc_crf_free(void *h) {
  orc_crf *m = (orc_crf *)h;
  if (!m) return;
  for (int k = 0; k < m->nterm; k++) {
    orc_lattice_free(m->term[k].lat);
    free(m->term[k].norm);
    free(m->term[k].cw);
  }
  free(m->term);
  free(m->unary);
  free(m->pQ);
  free(m);
}
void orc_crf_set_unary(void *h, const float *u_LxN) {
  orc_crf *m = (orc_crf *)h;
  free(m->unary);
  m->unary = orc_to_pixel_major(u_LxN, m->L, m->N);
}
/* LogisticUnaryEnergy::get = L*f (src/unary.cpp:50-52); summation order of Eigen's product is not
   restated -- this entry is compared with a tolerance, see tests. */
void orc_crf_set_unary_logistic(void *h, const float *Lm_LxF, int F, const float *f_FxN) {
  orc_crf *m = (orc_crf *)h;
  free(m->unary);
  m->unary = (float *)malloc(sizeof(float) * (size_t)m->L * m->N + 4);
  for (int i = 0; i < m->N; i++)
    for (int l = 0; l < m->L; l++) {
      float acc = 0.0f;
      for (int k = 0; k < F; k++) acc = acc + Lm_LxF[(size_t)l * F + k] * f_FxN[(size_t)k * m->N + i];
      m->unary[(size_t)i * m->L + l] = acc;
    }
}

/* DenseKernel::initLattice (src/pairwise.cpp:40-62) */
static float *orc_kernel_norm(orc_lattice *lat, int ntype) {
  const int N = lat->N;
  float *ones = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  float *nrm = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  for (int i = 0; i < N; i++) ones[i] = 1.0f;
  orc_lattice_filter(lat, nrm, ones, 1, 0);
  free(ones);
  if (ntype == ORC_NO_NORM) {
    float mean = 0.0f;
    for (int i = 0; i < N; i++) mean = mean + nrm[i]; /* sequential float sum, as written (:47-49) */
    mean = (float)N / mean;
    for (int i = 0; i < N; i++) nrm[i] = mean;
  } else if (ntype == ORC_NORM_SYMMETRIC) {
    for (int i = 0; i < N; i++) nrm[i] = (float)(1.0 / sqrt((double)nrm[i] + 1e-20));
  } else {
    for (int i = 0; i < N; i++) nrm[i] = (float)(1.0 / ((double)nrm[i] + 1e-20));
  }
  return nrm;
}

void orc_crf_add_pairwise(void *h, const float *feat_dxN, int d, int compat_kind, const float *compat,
                          int ktype, int ntype) {
  (void)ktype; /* KernelType only shapes the learnable parameters; inference ignores it (pairwise.cpp:116-121) */
  orc_crf *m = (orc_crf *)h;
  m->term = (orc_term *)realloc(m->term, sizeof(orc_term) * (size_t)(m->nterm + 1));
  orc_term *t = &m->term[m->nterm++];
  t->lat = (orc_lattice *)orc_lattice_build(feat_dxN, d, m->N);
  t->ntype = ntype;
  t->norm = orc_kernel_norm(t->lat, ntype);
  t->ckind = compat_kind;
  const int L = m->L;
  if (compat_kind == ORC_POTTS) {
    t->cw = (float *)malloc(sizeof(float));
    t->cw[0] = compat[0];
  } else if (compat_kind == ORC_DIAG) {
    t->cw = (float *)malloc(sizeof(float) * L);
    memcpy(t->cw, compat, sizeof(float) * L);
  } else {
    /* MatrixCompatibility ctor: 0.5*(m + m^T)  (src/labelcompatibility.cpp:79) */
    t->cw = (float *)malloc(sizeof(float) * (size_t)L * L);
    for (int a = 0; a < L; a++)
      for (int b = 0; b < L; b++) {
        float s = compat[(size_t)a * L + b] + compat[(size_t)b * L + a];
        t->cw[(size_t)a * L + b] = 0.5f * s;
      }
  }
}

/* DenseCRF2D feature builders (src/densecrf.cpp:61-81): float division of int coordinates / bytes */
int orc_crf_add_gaussian(void *h, float sx, float sy, int compat_kind, const float *compat, int ktype,
                         int ntype) {
  orc_crf *m = (orc_crf *)h;
  if (!m->W) return -1;
  float *f = (float *)malloc(sizeof(float) * 2 * (size_t)m->N);
  for (int y = 0; y < m->H; y++)
    for (int x = 0; x < m->W; x++) {
      f[(size_t)y * m->W + x] = (float)x / sx;
      f[(size_t)m->N + (size_t)y * m->W + x] = (float)y / sy;
    }
  orc_crf_add_pairwise(h, f, 2, compat_kind, compat, ktype, ntype);
  free(f);
  return 0;
}
int orc_crf_add_bilateral(void *h, float sx, float sy, float sr, float sg, float sb, const uint8_t *im,
                          int compat_kind, const float *compat, int ktype, int ntype) {
  orc_crf *m = (orc_crf *)h;
  if (!m->W) return -1;
  const size_t N = (size_t)m->N;
  float *f = (float *)malloc(sizeof(float) * 5 * N);
  for (int y = 0; y < m->H; y++)
    for (int x = 0; x < m->W; x++) {
      const size_t i = (size_t)y * m->W + x;
      f[i] = (float)x / sx;
      f[N + i] = (float)y / sy;
      f[2 * N + i] = (float)im[i * 3 + 0] / sr;
      f[3 * N + i] = (float)im[i * 3 + 1] / sg;
      f[4 * N + i] = (float)im[i * 3 + 2] / sb;
    }
  orc_crf_add_pairwise(h, f, 5, compat_kind, compat, ktype, ntype);
  free(f);
  return 0;
}

/* PairwisePotential::apply = DenseKernel::filter then compatibility (pairwise.cpp:63-80,173-178;
   labelcompatibility.cpp:46-48,66-69,82-84).  out and Q are NxL pixel-major. */
static void orc_term_apply(const orc_crf *m, const orc_term *t, float *out, const float *Q) {
  const int N = m->N, L = m->L;
  const int pre = (t->ntype == ORC_NORM_SYMMETRIC || t->ntype == ORC_NORM_BEFORE);
  const int post = (t->ntype == ORC_NORM_SYMMETRIC || t->ntype == ORC_NORM_AFTER);
  for (int i = 0; i < N; i++)
    for (int l = 0; l < L; l++) out[(size_t)i * L + l] = pre ? Q[(size_t)i * L + l] * t->norm[i] : Q[(size_t)i * L + l];
  orc_lattice_filter(t->lat, out, out, L, 0);
  if (post)
    for (int i = 0; i < N; i++)
      for (int l = 0; l < L; l++) out[(size_t)i * L + l] = out[(size_t)i * L + l] * t->norm[i];
  if (t->ckind == ORC_POTTS) {
    const float nw = -t->cw[0];
    for (size_t e = 0; e < (size_t)N * L; e++) out[e] = nw * out[e];
  } else if (t->ckind == ORC_DIAG) {
    for (int i = 0; i < N; i++)
      for (int l = 0; l < L; l++) out[(size_t)i * L + l] = t->cw[l] * out[(size_t)i * L + l];
  } else {
    /* out = W*Q through Eigen 3.3.9's GEBP micro-kernel (SSE2: mr=8, nr=4, unfused madd).  Every
       output accumulates over k in ascending order, except the L%4 left-over rows (see below). */
    float *col = (float *)malloc(sizeof(float) * L);
    const int Lal = (L / 4) * 4, Nal = (N / 4) * 4;
    for (int i = 0; i < N; i++) {
      float *o = out + (size_t)i * L;
      memcpy(col, o, sizeof(float) * L);
      for (int l = 0; l < L; l++) {
        const float *wr = t->cw + (size_t)l * L;
        float acc;
        if (l >= Lal && i < Nal) {
          /* left-over rows (L%4) inside whole 4-column panels: four interleaved partial sums over
             k, combined as (c0+c1)+(c2+c3), then the k tail  (GeneralBlockPanelKernel.h:1521-1580) */
          float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
          const int k4 = (L / 4) * 4;
          for (int k = 0; k < k4; k += 4) {
            float p0 = wr[k] * col[k], p1 = wr[k + 1] * col[k + 1];
            float p2 = wr[k + 2] * col[k + 2], p3 = wr[k + 3] * col[k + 3];
            c0 = c0 + p0;
            c1 = c1 + p1;
            c2 = c2 + p2;
            c3 = c3 + p3;
          }
          float s01 = c0 + c1, s23 = c2 + c3;
          acc = s01 + s23;
          for (int k = k4; k < L; k++) {
            float p = wr[k] * col[k];
            acc = acc + p;
          }
        } else {
          acc = 0.0f;
          for (int k = 0; k < L; k++) {
            float p = wr[k] * col[k];
            acc = acc + p;
          }
        }
        o[l] = acc;
      }
    }
    free(col);
  }
}

void orc_crf_term_apply(void *h, int k, float *out_NxL, const float *Q_NxL) {
  orc_crf *m = (orc_crf *)h;
  orc_term_apply(m, &m->term[k], out_NxL, Q_NxL);
}
void orc_crf_term_norm(void *h, int k, float *norm_N) {
  orc_crf *m = (orc_crf *)h;
  memcpy(norm_N, m->term[k].norm, sizeof(float) * (size_t)m->N);
}
void *orc_crf_term_lattice(void *h, int k) { return ((orc_crf *)h)->term[k].lat; }

/* one mean-field update (the body of src/densecrf.cpp:122-129 and :187-200) */
static void orc_step(const orc_crf *m, float *Q, int from_zero) {
  const size_t n = (size_t)m->N * m->L;
  float *t1 = (float *)malloc(sizeof(float) * (n + 1));
  float *t2 = (float *)malloc(sizeof(float) * (n + 1));
  if (from_zero) { /* stepInference: tmp1 = 0; tmp1 -= unary (:188-191) */
    for (size_t e = 0; e < n; e++) t1[e] = 0.0f;
    if (m->unary)
      for (size_t e = 0; e < n; e++) t1[e] = t1[e] - m->unary[e];
  } else { /* inference: tmp1 = -unary (:123); unary is all-zero when none was set (:117) */
    for (size_t e = 0; e < n; e++) t1[e] = m->unary ? -m->unary[e] : -0.0f;
  }
  for (int k = 0; k < m->nterm; k++) {
    orc_term_apply(m, &m->term[k], t2, Q);
    for (size_t e = 0; e < n; e++) t1[e] = t1[e] - t2[e];
  }
  orc_exp_and_normalize(Q, t1, m->L, m->N);
  free(t1);
  free(t2);
}

static void orc_q0(const orc_crf *m, float *Q) {
  const size_t n = (size_t)m->N * m->L;
  float *t = (float *)malloc(sizeof(float) * (n + 1));
  for (size_t e = 0; e < n; e++) t[e] = m->unary ? -m->unary[e] : -0.0f;
  orc_exp_and_normalize(Q, t, m->L, m->N);
  free(t);
}

/* DenseCRF::inference (src/densecrf.cpp:115-131) */
void orc_crf_inference(void *h, int n_iter, float *Q_NxL) {
  orc_crf *m = (orc_crf *)h;
  orc_q0(m, Q_NxL);
  for (int it = 0; it < n_iter; it++) orc_step(m, Q_NxL, 0);
}
/* DenseCRF::startInference (:178-186): without a unary Q stays all-zero */
void orc_crf_start(void *h, float *Q_NxL) {
  orc_crf *m = (orc_crf *)h;
  if (m->unary)
    orc_q0(m, Q_NxL);
  else
    memset(Q_NxL, 0, sizeof(float) * (size_t)m->N * m->L);
}
/* DenseCRF::stepInference (:187-201) */
void orc_crf_step(void *h, float *Q_NxL) { orc_step((orc_crf *)h, Q_NxL, 1); }

/* The stepwise protocol as pydensecrf's wrapper drives it (densecrf.pyx:94-98): Q = startInference() once, then
   stepInference on the same matrix, no copies in or out.  Used as the timed CPU arm when oracle/_ref is absent. */
void orc_crf_persist_start(void *h) {
  orc_crf *m = (orc_crf *)h;
  free(m->pQ);
  m->pQ = (float *)malloc(sizeof(float) * ((size_t)m->N * m->L + 4));
  orc_crf_start(h, m->pQ);
}
void orc_crf_persist_step(void *h) { orc_step((orc_crf *)h, ((orc_crf *)h)->pQ, 1); }
void orc_crf_persist_get(void *h, float *Q_NxL) {
  orc_crf *m = (orc_crf *)h;
  memcpy(Q_NxL, m->pQ, sizeof(float) * (size_t)m->N * m->L);
}

/* DenseCRF::klDivergence (:214-235): float products, double accumulator for the first two terms,
   Eigen's float vectorised sum for each pairwise term. */
double orc_crf_kl(void *h, const float *Q) {
  orc_crf *m = (orc_crf *)h;
  const size_t n = (size_t)m->N * m->L;
  double kl = 0.0;
  for (size_t e = 0; e < n; e++) {
    float q = Q[e];
    float c = (q > 1e-20f) ? q : 1e-20f;
    kl += (double)q * log((double)c); /* unqualified log(float) binds to ::log(double) here */
  }
  if (m->unary)
    for (size_t e = 0; e < n; e++) {
      float p = m->unary[e] * Q[e];
      kl += (double)p;
    }
  float *tmp = (float *)malloc(sizeof(float) * (n + 4));
  for (int k = 0; k < m->nterm; k++) {
    orc_term_apply(m, &m->term[k], tmp, Q);
    for (size_t e = 0; e < n; e++) tmp[e] = Q[e] * tmp[e];
    kl += (double)orc_eigen_sum(tmp, (long)n);
  }
  free(tmp);
  return kl;
}

const char *orc_describe(void) { return "C restatement of the reference hot path (oracle/dcrf_oracle.c)"; }
