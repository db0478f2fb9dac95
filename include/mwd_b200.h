/*
 * mwd_b200.h -- C ABI of the B200-native EM hot path for MultimodalWordDiscovery's
 * HMM / HMM-DNN word discoverers.
 *
 * The reference is pure Python/NumPy and has no FFI; its "operator boundary" is the method
 * surface of its word-discoverer classes.  Each entry point below replaces the per-caption Python
 * loop behind one (group of) reference method(s); the file:line it stands in for is cited.  A
 * maintainer of the reference binds these with ctypes; the package
 * `multimodalworddiscovery_b200` does exactly that.
 *
 * Conventions
 *   - plain pointers and sizes only; every `*_d` / `const double*` etc. argument marked [dev] is a
 *     DEVICE pointer on the current CUDA device, [host] is a host pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising unless stated otherwise;
 *   - return value 0 = success, non-zero = failure; `mwd_last_error()` gives the message
 *     (thread-local); nothing is thrown across the boundary;
 *   - all probabilities / counts are IEEE float64 in the raw probability domain, exactly as in
 *     the reference (EPS = 1e-50 floors included); region features are float32 or float64.
 *
 * Corpus layout ("packed pairs").  Pairs (caption, image) are sorted by (n regions, T phones) and
 * stored CSR-style:
 *     region_off[N+1] (int32)  -> rows of feats[R][D]   (row-major, D contiguous)
 *     phone_off [N+1] (int32)  -> entries of phones[Ttot] (int32 phone ids in [0,P))
 * A "bucket" is a contiguous range of pairs with the same n: pairs [bucket_lo[b], bucket_lo[b+1])
 * have n == bucket_n[b].
 *
 * Parameter layout
 *     init  [(MWD_NMAX+1)][MWD_NMAX]            init[m][i]      (reference: self.init[m][i])
 *     trans [(MWD_NMAX+1)][MWD_NMAX*MWD_NMAX]   trans[m][i*m+j] (reference: self.trans[m][i][j])
 *     obsT  [P][K]                              obsT[p][k] = self.obs[k][p]  (transposed)
 *     W     [K][D+1]   (linear)   |   mus [K][D]  (gaussian)
 */
#ifndef MWD_B200_H
#define MWD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWD_NMAX 16          /* max regions (HMM states) per image                       */
#define MWD_KMAX 128         /* max concepts (nWords)                                     */
#define MWD_EPS 1e-50        /* hmm_dnn/image_phone_hmm_word_discoverer.py:11             */
#define MWD_MIXED_CONCEPT 1
#define MWD_MIXED_POSTERIOR 2
#define MWD_MIXED_GRAD 4
#define MWD_MIXED_RECURSION 8
#define MWD_INIT_STRIDE MWD_NMAX
#define MWD_TRANS_STRIDE (MWD_NMAX * MWD_NMAX)

const char* mwd_last_error(void);
int mwd_version(void);
/* sizeof() of the ABI structs as compiled (which: 0 mwd_geometry, 1 mwd_ik_problem,
 * 2 mwd_partial_sizes, 3 mwd_ik_mstep_args, 4 mwd_hmm_problem, 5 mwd_hmm_mstep_args) -- lets a
 * binding verify its struct mirrors.                                                        */
int mwd_abi_sizeof(int which);

/* device / launch geometry the host side needs to size workspaces ------------------------- */
typedef struct {
  int32_t sm_count;          /* multiprocessors of the current device                      */
  int32_t estep_grid;        /* persistent grid of mwd_ik_estep (= rows of the partial tables) */
  int32_t grad_splits;       /* row-splits of mwd_ik_posterior_grad                          */
} mwd_geometry;
int mwd_get_geometry(mwd_geometry* out);

/* packed corpus + parameters + workspaces of the (region i, concept k)-state model -------- */
typedef struct {
  /* corpus */
  int64_t n_pairs;           /* N (this rank's shard)                                       */
  int64_t n_regions;         /* R = region_off[N]                                           */
  int64_t n_phones_total;    /* Ttot = phone_off[N]                                         */
  int32_t feat_dim;          /* D                                                           */
  int32_t feat_is_f64;       /* 0: feats are float32, 1: float64                            */
  int32_t n_concepts;        /* K  (<= MWD_KMAX)                                            */
  int32_t n_phone_types;     /* P                                                           */
  int32_t t_max;             /* longest caption in the shard                                */
  int32_t n_buckets;
  const int32_t* bucket_n;   /* [host] n of each bucket                                     */
  const int64_t* bucket_lo;  /* [host] n_buckets+1 pair offsets                             */
  const int32_t* bucket_tmax;/* [host] longest caption of each bucket                       */
  const int32_t* region_off; /* [dev]                                                       */
  const int32_t* phone_off;  /* [dev]                                                       */
  const void*    feats;      /* [dev] R x D                                                 */
  const int32_t* phones;     /* [dev] Ttot                                                  */
  /* parameters */
  const double* init;        /* [dev] see layout above                                      */
  const double* trans;       /* [dev]                                                       */
  const double* obsT;        /* [dev] P x K                                                 */
  /* per-region work/outputs */
  double* pz;                /* [dev] R x K  image posterior p(z|v)                         */
  double* concept_counts;    /* [dev] R x K  reference: self.conceptCounts                  */
  /* per-pair outputs */
  double* pair_ll;           /* [dev] N      log(max(sum alpha_{T-1}, EPS))                 */
  double* concept_counts_a;  /* [dev] Ttot x K or NULL  reference: self.conceptCountsA      */
  /* count partials (one row per persistent CTA) and scratch */
  double* part_phone;        /* [dev] estep_grid x P x K                                    */
  double* part_init;         /* [dev] estep_grid x (MWD_NMAX+1) x MWD_NMAX                  */
  double* part_trans;        /* [dev] estep_grid x (MWD_NMAX+1) x MWD_NMAX*MWD_NMAX         */
  double* scratch;           /* [dev] checkpoint scratch, scratch_bytes long                */
  int64_t scratch_bytes;
  double* stats;             /* [dev] 4 * slot_off[N] doubles: per (pair, t) row statistics
                                (s_t, floor-sum, xi diagonal, r_t; n each) handed from the
                                recursion kernel to the count post-pass                      */
  const int64_t* slot_off;   /* [dev] N+1: slot_off[p] = sum_{q<p} T_q * n_q                */
  int32_t no_floor;          /* 0: reference EPS floors (all classes but one); 1: the un-floored
                                normalisers of ImageAudioGaussianHMMWordDiscoverer
                                (image_audio_gaussian_hmm_word_discoverer.py:369-371,414-415,449-451,
                                :629-631): likelihood, gamma and xi are divided by their raw sums   */
  int32_t mixed_precision;   /* 0: everything in float64, bit for bit the reference's arithmetic class (default).
                                Bits (MWD_MIXED_*) move the parts of the iteration that have NO EPS floor
                                off the FP64 pipe; they are validated against the float64 path at the
                                north-star tolerance (1e-5 on log-likelihood and tables), never bit-exact:
                                  MWD_MIXED_CONCEPT    updateConceptCounts chains in float32 (FFMA pipe)
                                  MWD_MIXED_POSTERIOR  softmaxLayer GEMM on tcgen05 (kind::tf32, operands split
                                                       hi + lo, fp32 TMEM accumulators, float64 recombination
                                                       + softmax), features through TMA; linear class, fp32
                                                       features (mwd_posterior_linear_tc)
                                  MWD_MIXED_GRAD       updateSoftmaxWeight GEMM likewise (mwd_ik_posterior_grad_tc)
                                  MWD_MIXED_RECURSION  forward / backward lattice in scaled float32 (one power-of-two
                                                       exponent per (pair, t)); every EPS floor is applied to the
                                                       scaled value with the scaled threshold, the statistics handed
                                                       to the count kernels and the phone counts stay float64
                                                       (csrc/ik_estep_warp32.cu)                             */
  int32_t* concept_alignment;/* [dev] Ttot or NULL: argmax_k conceptCountsA[t][k] (first index on ties, the
                                `concept_alignment` of printAlignment :628) written by mwd_ik_estep from
                                the column sums it forms anyway -- 4 bytes per phone instead of the
                                8K-byte conceptCountsA row                                          */
} mwd_ik_problem;

/* bytes of `scratch` mwd_ik_estep needs for this problem (depends on t_max, bucket_n, K) */
int64_t mwd_ik_scratch_bytes(const mwd_ik_problem* p);

/* softmaxLayer -- image_phone_hmm_word_discoverer.py:533-541:
 *   pz[r][k] = softmax_k( feats[r] . W[k][0:D] + W[k][D] )                                  */
int mwd_posterior_linear(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                         const double* W, int n_concepts, double* pz, void* stream);

/* softmaxLayer on the Blackwell tensor cores (tcgen05.mma kind::tf32 + TMA + TMEM), the MWD_MIXED_POSTERIOR path:
 * same result contract as mwd_posterior_linear to ~1e-6 relative (split-TF32 operands, see csrc/posterior_tc.cu);
 * fp32 features only.  w_split_scratch [dev]: mwd_posterior_tc_scratch_bytes(K, D) bytes, 16-byte aligned.
 * split_mode 0: both feature parts rounded to TF32 in shared memory; 1: the high part is the raw fp32 word
 * (relies on the tensor core ignoring the low 13 mantissa bits).  mwd_posterior_tc_supported() tells whether a
 * shape can take this path (fp32 features, D % 4 == 0, D >= 32, K <= MWD_KMAX).                              */
int64_t mwd_posterior_tc_scratch_bytes(int n_concepts, int feat_dim);
int mwd_posterior_tc_supported(int feat_is_f64, int feat_dim, int n_concepts);
int mwd_posterior_linear_tc(const float* feats, int64_t n_regions, int feat_dim, const double* W,
                            int n_concepts, double* pz, void* w_split_scratch, int split_mode, void* stream);

/* updateSoftmaxWeight GEMM on the Blackwell tensor cores, the MWD_MIXED_GRAD path (csrc/posterior_grad_tc.cu):
 *   grad[k][d] = sum_r (concept_counts - pz)[r][k] * [feats,1][r][d]   -- image_phone_hmm_word_discoverer.py:475-488
 * split-TF32 operands, fp32 TMEM accumulation over at most 512 rows, float64 per-CTA partial tables summed in fixed
 * order (deterministic).  partials [dev]: mwd_posterior_grad_tc_partials_len(K, D) doubles.  _partial with
 * accumulate == 0 zeroes the partial tables first; accumulate != 0 adds another chunk of the shard.  _finish writes
 * grad (K x (D+1)).  Uses p->feats (fp32), p->concept_counts, p->pz, p->n_regions, p->feat_dim, p->n_concepts.     */
int mwd_posterior_grad_tc_supported(int feat_is_f64, int feat_dim, int n_concepts);
int64_t mwd_posterior_grad_tc_partials_len(int n_concepts, int feat_dim);
int mwd_ik_posterior_grad_tc_partial(const mwd_ik_problem* p, double* partials, int accumulate, int split_mode,
                                     void* stream);
int mwd_posterior_grad_tc_finish(int n_concepts, int feat_dim, const double* partials, double* grad, void* stream);

/* Diagnostic (tools/umma_layout_probe.py): n_mma tcgen05.mma kind::tf32 on caller-supplied shared-memory images and
 * matrix descriptors (start-address field zero), the 128 x n_cols fp32 accumulator dumped to out.                */
int mwd_umma_probe(const void* a_img, int a_words, const void* b_img, int b_words, uint64_t adesc, uint64_t bdesc,
                   uint32_t idesc, int n_cols, int n_mma, uint32_t a_step, uint32_t b_step, float* out, void* stream);

/* softmaxLayer -- image_phone_gaussian_hmm_word_discoverer.py:501-510:
 *   pz[r][k] = softmax_k( -||feats[r] - mus[k]||^2 / width )
 * w_scratch [dev] K x (D+1) receives the expanded weights [2 mu/width , -||mu||^2/width]      */
int mwd_posterior_gaussian(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                           const double* mus, double width, int n_concepts, double* w_scratch,
                           double* pz, void* stream);

/* ---- two-layer posterior of ImagePhoneHMMDNNWordDiscoverer -----------------------
 * hiddenLayer -- hmm_dnn/image_phone_hmm_dnn_word_discoverer.py:573-579:
 *   hidden[r][h] = relu( feats[r] . V[h][0:D] + V[h][D] )             (V: H x (D+1))
 * the image posterior is then mwd_posterior_linear(hidden (float64, feat_dim = H), W: K x (H+1)). */
int mwd_hidden_relu(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                    const double* V, int hidden_dim, double* hidden, void* stream);
/* ReLU back-propagation of updateNeuralNetWeights (:510-511,526):
 *   eps[r][h] = (hidden[r][h] > 0) * sum_k (concept_counts - pz)[r][k] * W[k][h]              */
int mwd_backprop_hidden(const double* concept_counts, const double* pz, const double* W,
                        const double* hidden, int64_t n_regions, int n_concepts, int hidden_dim,
                        double* eps, void* stream);
/* grad[m][d] = sum_r (delta - minus)[r][m] * [feats[r], 1][d]   (n_rows_out x (D+1), UNscaled);
 * `minus` may be NULL.  The two weight gradients of :522-526 are two calls of this
 * (delta = concept_counts, minus = pz over the hidden activations; delta = eps over the features).
 * grad_partials [dev]: mwd_outer_grad_partials_len(n_rows_out, D) doubles of scratch (row splits
 * x min(n_rows_out,128) x (D+1); narrow feature matrices get more row splits).                 */
int64_t mwd_outer_grad_partials_len(int n_rows_out, int feat_dim);
int mwd_outer_grad(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                   const double* delta, const double* minus, int n_rows_out, double* grad_partials,
                   double* grad, void* stream);
/* param = (1 - momentum) * param + lr * scale * grad   (:527-528, scale = 1/N)                 */
int mwd_sgd_update(double* param, const double* grad, int64_t elems, double scale, double lr,
                   double momentum, void* stream);

/* ---- dense-emission classes: ImageAudioHMMWordDiscoverer -------------------------
 * hmm_dnn/image_audio_hmm_word_discoverer.py replaces the discrete emission obs[:, x_t] by
 *     E[t][k] = sum_ph phoneProbs[k][ph] * p(ph | a_t)          (:286-288, :320-322, :384-386)
 * with p(ph | a_t) = softmaxLayerA(a_t) (:549-554) = mwd_posterior_linear over the audio frames.
 * The recursion / concept / decode entry points run unchanged on
 *     obsT = emis (n_frames x K),  phones[f] = f (identity),  n_phone_types = n_frames,
 *     part_phone = NULL (no phone table: mwd_ik_estep then only fills concept_counts_a and
 *     mwd_ik_reduce_counts zeroes the phone block of `counts`).
 * mwd_dense_emission:      emis[f][k] from frame_post (n_frames x n_phones) and phone_probs_t
 *                          (n_phones x K, the obsT layout of phoneProbs); a GEMM on the FP64 tensor
 *                          path.  v_scratch [dev]: K x (n_phones+1) doubles.
 * mwd_concept_phone_counts: updateConceptPhoneCounts (:486-493) summed over the frames, i.e. the
 *                          phoneCounts of trainUsingEM :230-231 in the obsT layout:
 *     counts_t[ph][k] = sum_f cA[f][k] * frame_post[f][ph] / (sum_k cA[f] * sum_ph frame_post[f])
 *                          concept_counts_a is NORMALISED IN PLACE (rows divided by the bracket), then
 *                          reduced by the outer-product GEMM with deterministic row-split partials.
 *                          partials [dev]: mwd_concept_phone_partials_len() doubles of scratch.      */
int mwd_dense_emission(const double* frame_post, const double* phone_probs_t, int64_t n_frames,
                       int n_phones, int n_concepts, double* v_scratch, double* emis, void* stream);
int64_t mwd_concept_phone_partials_len(int n_concepts, int n_phones);
int mwd_concept_phone_counts(double* concept_counts_a, const double* frame_post, int64_t n_frames,
                             int n_concepts, int n_phones, double* partials, double* counts_t,
                             void* stream);

/* forward + backward + updateInitialCounts + updateTransitionCounts + updateStateCounts +
 * computeAvgLogLikelihood -- image_phone_hmm_word_discoverer.py:276-433, 523-531, and the
 * phoneCounts / conceptCountsA accumulation of trainUsingEM :230-235.
 * Reads p->pz; writes pair_ll, (concept_counts_a), and ACCUMULATES into part_phone / part_init /
 * part_trans (the caller zeroes them at the start of an EM iteration).  part_init/part_trans hold one table
 * per n: layout [estep_grid][MWD_NMAX+1][...] -- see mwd_ik_partial_sizes.                   */
int mwd_ik_estep(const mwd_ik_problem* p, void* stream);

/* computeAvgLogLikelihood alone -- :523-531: forward sweep only, writes p->pair_ll.          */
int mwd_ik_loglik(const mwd_ik_problem* p, void* stream);

typedef struct {
  int64_t phone_elems;       /* estep_grid * P * K                                          */
  int64_t init_elems;        /* estep_grid * (MWD_NMAX+1) * MWD_NMAX                         */
  int64_t trans_elems;       /* estep_grid * (MWD_NMAX+1) * MWD_NMAX * MWD_NMAX              */
} mwd_partial_sizes;
int mwd_ik_partial_sizes(int n_concepts, int n_phone_types, mwd_partial_sizes* out);

/* updateConceptCounts -- image_phone_hmm_word_discoverer.py:443-465.  Reads p->pz, writes
 * p->concept_counts.                                                                        */
int mwd_ik_concept_counts(const mwd_ik_problem* p, void* stream);

/* Deterministic second-level reduction of the per-CTA partials (fixed order, no atomics):
 *   counts = [ phoneC (P x K, transposed) | initC ((NMAX+1) x NMAX) | transC ((NMAX+1) x NMAX^2) | sum LL ]
 * `counts` [dev] has mwd_ik_counts_len(K,P) doubles; this is the buffer that is all-reduced
 * across GPUs before the M-step.  The partial tables are CONSUMED (the first reduction level
 * parks its sums in place): zero them before the next E-step, call this once per iteration.   */
int64_t mwd_ik_counts_len(int n_concepts, int n_phone_types);
int mwd_ik_reduce_counts(const mwd_ik_problem* p, double* counts, void* stream);

/* Gradient of the image-posterior parameters -- updateSoftmaxWeight,
 * image_phone_hmm_word_discoverer.py:475-488 (linear) / gaussian :488-499:
 *   grad[k][d] = sum_r (concept_counts - pz)[r][k] * [feats[r], 1][d]     (K x (D+1), UNscaled)
 * grad_partials [dev] : mwd_outer_grad_partials_len(K, D) doubles;  grad [dev] : K x (D+1).    */
int mwd_ik_posterior_grad(const mwd_ik_problem* p, double* grad_partials, double* grad,
                          void* stream);
/* Streaming form for corpora processed in chunks (host-resident or larger than HBM): add one
 * chunk's rows into the partials (accumulate = 0 for the first chunk), then reduce once.     */
int mwd_ik_posterior_grad_partial(const mwd_ik_problem* p, double* grad_partials, int accumulate,
                                  void* stream);
int mwd_ik_posterior_grad_finish(int n_concepts, int feat_dim, const double* grad_partials,
                                 double* grad, void* stream);

/* M-step of trainUsingEM -- image_phone_hmm_word_discoverer.py:238-258 (gaussian :238-264).
 * Consumes the (globally reduced) `counts` and `grad`, updates init / trans / obsT and W or mus
 * in place.  lens[n_lens] [host] are the distinct n of the WHOLE corpus (reference: self.lenProb
 * keys); toeplitz = (n_lens >= 6) pooling of :399-413 is applied here (it is linear, so it
 * commutes with the sums over t and over pairs).  n_pairs_global is len(self.vCorpus).       */
#define MWD_MSTEP_FLOOR_TABLES 1   /* EPS-floor init/trans counts (gaussian and two-layer classes)   */
#define MWD_MSTEP_NO_POSTERIOR 2  /* leave posterior_param alone (two-layer: mwd_sgd_update instead)*/
#define MWD_MSTEP_FREEZE_TRANS 4  /* trainUsingEM(freezeTransition=True) of the two-layer class     */
#define MWD_MSTEP_NO_FLOORS 8     /* un-floored init / trans / phone-table normalisers even when gaussian
                                     (image_audio_gaussian_hmm_word_discoverer.py:241-260)           */
typedef struct {
  int32_t gaussian;          /* 0 linear (W), 1 gaussian (mus; implies FLOOR_TABLES)        */
  int32_t n_concepts, n_phone_types, feat_dim;
  int32_t n_lens;
  int32_t flags;             /* MWD_MSTEP_* bits                                            */
  const int32_t* lens;       /* [host]                                                      */
  int32_t toeplitz;
  int64_t n_pairs_global;
  double lr, momentum, width;
  const double* counts;      /* [dev]                                                       */
  const double* grad;        /* [dev] K x (D+1)                                             */
  double* init;              /* [dev] in/out                                                */
  double* trans;             /* [dev] in/out                                                */
  double* obsT;              /* [dev] out                                                   */
  double* posterior_param;   /* [dev] in/out  W (K x (D+1)) or mus (K x D)                  */
} mwd_ik_mstep_args;
int mwd_ik_mstep(const mwd_ik_mstep_args* a, void* stream);

/* align + cluster of printAlignment -- image_phone_hmm_word_discoverer.py:543-597, 620-648.
 * Reads p->pz.
 *   alignment        [dev] Ttot   int32  Viterbi region index per phone (bit-exact tie rules);
 *                                        INPUT instead when given_alignment != 0 (cluster() only)
 *   align_probs      [dev] sum_p T_p*n_p doubles or NULL  (offset of pair p = ap_off[p])
 *   ap_off           [dev] N+1 int64 (required iff align_probs != NULL)
 *   image_concepts   [dev] R      int32  cluster() argmax
 *   cluster_scores   [dev] R x K doubles or NULL  cluster() scores
 *   floor_norm: bit 0 = floored alignProbs normaliser (gaussian :583, two-layer :622);
 *               bit 1 = do NOT floor the Viterbi scores (two-layer class, :612)              */
int mwd_ik_decode(const mwd_ik_problem* p, int floor_norm, int given_alignment, int32_t* alignment,
                  double* align_probs, const int64_t* ap_off, int32_t* image_concepts,
                  double* cluster_scores, void* stream);

/* printAlignment's two files -- image_phone_hmm_word_discoverer.py:620-648: HOST-side
 * writer, byte-for-byte `'%d ' per phone + blank line` (.txt) and
 * `json.dump(aligns, f, indent=4, sort_keys=True)` (.json, floats as float.__repr__) straight from the
 * flat arrays of mwd_ik_decode.  All pointers [host], pairs in corpus order; phone_off / region_off:
 * n_pairs+1 offsets; align_probs: (T x n) per pair at ap_off[p]; concept_alignment (Ttot),
 * concept_probs (R x K, gaussian :651) and cluster_probs (R x K, two-layer) may be NULL -> key omitted. */
int mwd_write_alignment_files(const char* txt_path, const char* json_path, int64_t n_pairs,
                              const int64_t* phone_off, const int64_t* region_off, const int32_t* alignment,
                              const int32_t* image_concepts, const int32_t* concept_alignment,
                              const double* align_probs, const int64_t* ap_off, const double* concept_probs,
                              const double* cluster_probs, int n_concepts, int is_phoneme);
/* float.__repr__(v) into buf (NUL-terminated); returns the length or -1 (test hook of the writer) */
int mwd_format_float_repr(double v, char* buf, int buf_len);

/* Stream-ordered utilities of the host mirror (no reference counterpart: the reference zeroes its
 * count dicts in Python, trainUsingEM :209-213, and sums the log-likelihood in
 * computeAvgLogLikelihood :523-531; the rank combination belongs to the multi-GPU sharding).
 *   mwd_fill_f64    p[0:n] = value (cudaMemsetAsync when value == 0)
 *   mwd_sum_f64     out[0] = fixed-shape deterministic sum of x[0:n]; scratch256: >= 256 doubles
 *   mwd_rank_reduce out[e] = sum_r gathered[r][e] in rank order (log_domain != 0: logsumexp over r,
 *                   except entry ll_index, which stays a plain sum; pass -1 for none)             */
int mwd_fill_f64(double* p, int64_t n, double value, void* stream);
int mwd_sum_f64(const double* x, int64_t n, double* scratch256, double* out, void* stream);
int mwd_rank_reduce(const double* gathered, int world, int64_t n, int log_domain, int64_t ll_index,
                    double* out, void* stream);

/* argmax_k rows[t][k] (first index on ties, NaN-aware like np.argmax) -- printAlignment :628 */
int mwd_argmax_rows(const double* rows, int64_t n_rows, int n_cols, int32_t* out, void* stream);

/* Dense forward / backward of ONE pair -- forward :276-304 / backward :314-335.
 * pz_pair [dev] n x K, phones_pair [dev] T; out [dev] T x n x K.                            */
int mwd_ik_forward_dense(const double* pz_pair, const int32_t* phones_pair, int T, int n, int K,
                         const double* init, const double* trans, const double* obsT,
                         double* out, void* stream);
int mwd_ik_backward_dense(const double* pz_pair, const int32_t* phones_pair, int T, int n, int K,
                          const double* trans, const double* obsT, double* out, void* stream);

/* ============================================================================================
 * Plain-state HMM word discoverers of hmm/  (states = concept tokens of the caption)
 *   prob domain: hmm/hmm_word_discoverer.py        HMMWordDiscoverer
 *   log  domain: hmm/audio_hmm_word_discoverer.py  AudioHMMWordDiscoverer (NULL state first;
 *                hmm/hmm_word_discoverer_logscale.py is the same code)
 * The dict-of-dict obs[tw][fw] is a dense (Vt x Vf) table, NaN = pair absent from the dict.
 * Pairs are sorted by (n states, T) and CSR-packed like the (i,k) model:
 *   tgt_off[N+1] -> tgt[] concept ids (the states), src_off[N+1] -> src[] phone ids.
 * Slot (pair p, t, i) of the per-pair posterior buffer lives at slot_off[p] + t*n_p + i.
 * ============================================================================================ */
typedef struct {
  int64_t n_pairs;
  int64_t n_slots;            /* sum_p T_p * n_p                                             */
  int32_t n_tgt_types;        /* Vt                                                          */
  int32_t n_src_types;        /* Vf                                                          */
  int32_t t_max;
  int32_t log_domain;         /* 0: HMMWordDiscoverer, 1: AudioHMMWordDiscoverer             */
  int32_t n_buckets;
  int32_t reserved;
  const int32_t* bucket_n;    /* [host]                                                      */
  const int64_t* bucket_lo;   /* [host] n_buckets+1                                          */
  const int32_t* bucket_tmax; /* [host]                                                      */
  const int32_t* tgt_off;     /* [dev] N+1                                                   */
  const int32_t* tgt;         /* [dev]                                                       */
  const int32_t* src_off;     /* [dev] N+1                                                   */
  const int32_t* src;         /* [dev]                                                       */
  const int64_t* slot_off;    /* [dev] N+1                                                   */
  const double* init;         /* [dev] (NMAX+1) x NMAX        (log values when log_domain)   */
  const double* trans;        /* [dev] (NMAX+1) x NMAX*NMAX                                  */
  const double* obs;          /* [dev] Vt x Vf, NaN = absent                                 */
  double* pair_ll;            /* [dev] N   log p(f | e) per pair                             */
  double* post;               /* [dev] n_slots  state posteriors (prob) / normalised log
                                 posteriors (log) -- input of the postings reduction         */
  double* part_init;          /* [dev] hmm_warps x (NMAX+1) x NMAX                           */
  double* part_trans;         /* [dev] hmm_warps x (NMAX+1) x NMAX*NMAX                      */
  double* alpha_out;          /* [dev] n_slots or NULL: forward()  values                    */
  double* beta_out;           /* [dev] n_slots or NULL: backward() values                    */
  const double* emis;         /* [dev] n_slots or NULL: dense log emissions lb[t][j] (segment-
                                 embedding model); when set, `obs` is not read                */
  int64_t n_src_rows;         /* src_off[N] (segment model: rows of the embedding matrix)     */
  const int32_t* row_pair;    /* [dev] n_src_rows: pair of each source row (segment model)    */
  const int32_t* slot_row;    /* [dev] n_slots: source row of each slot       (segment model) */
} mwd_hmm_problem;

/* number of persistent warps (= rows of part_init / part_trans)                              */
int mwd_hmm_warps(void);

/* forward + backward + updateInitialCounts + updateTransitionCounts + state posteriors:
 * hmm_word_discoverer.py:110-203 / audio_hmm_word_discoverer.py:148-254 (incl. its quirks:
 * un-normalised init counts, transition counts from the last t only).  Accumulates into
 * part_init / part_trans (zero / -inf them first), writes pair_ll and post.
 * Transient device scratch (the alpha lattices of the resident warps; in mwd_hmm_reduce and
 * mwd_hmm_gauss_stats the split-reduction partials) comes from a stream-ordered memory pool the
 * library owns per device and is returned to it, in stream order, before the call returns.   */
int mwd_hmm_estep(const mwd_hmm_problem* p, void* stream);

/* Deterministic reduction of the partials and of the posteriors through a static postings
 * index (post_idx sorted by table entry, post_off[Vt*Vf+1]):
 *   counts = [ obsC (Vt x Vf) | initC ((NMAX+1) x NMAX) | transC ((NMAX+1) x NMAX^2) | sum LL ]
 * sums in the prob domain, log-sum-exp in the log domain.  p->n_slots must equal
 * post_off[Vt*Vf] (it bounds the list of entries that take the split reduction).             */
int64_t mwd_hmm_counts_len(int n_tgt_types, int n_src_types);
int mwd_hmm_reduce(const mwd_hmm_problem* p, const int64_t* post_idx, const int64_t* post_off,
                   double* counts, void* stream);

/* M-step: hmm_word_discoverer.py:275-296 (Toeplitz pooling applied here, it is linear) /
 * audio_hmm_word_discoverer.py:354-389 (counts first merged into the running log accumulators
 * `acc`, which persist over epochs like the reference's count lists).                        */
typedef struct {
  int32_t log_domain, n_tgt_types, n_src_types, n_lens;   /* n_src_types == 0: skip the obs table */
  const int32_t* lens;        /* [host]                                                      */
  const double* counts;       /* [dev]                                                       */
  double* acc;                /* [dev] running log accumulators (log domain only), same layout*/
  double* init;               /* [dev] in/out                                                */
  double* trans;              /* [dev] in/out                                                */
  double* obs;                /* [dev] in/out (NaN entries stay NaN)                         */
} mwd_hmm_mstep_args;
int mwd_hmm_mstep(const mwd_hmm_mstep_args* a, void* stream);

/* align: hmm_word_discoverer.py:301-329 / audio_hmm_word_discoverer.py:396-427.
 *   alignment   [dev] sum_p T_p int32
 *   align_probs [dev] sum_p (T_p - 1) * n_p doubles or NULL, pair p at ap_off[p]             */
int mwd_hmm_align(const mwd_hmm_problem* p, double unk_prob, int32_t* alignment, double* align_probs,
                  const int64_t* ap_off, void* stream);

/* ---- segment-embedding HMM (hmm/audio_segembed_hmm_word_discoverer.py + the emission model it
 * was written against, smt/audio_gmm_word_discoverer.py:53-61,395-401): source "tokens" are
 * rows of emb[S][D]; state j of pair p emits x_t with
 *     lb[t][j] = LSE_m( lprior[w][m] + log N(x_t; means[w][m], diag var[w][m]) ),  w = tgt[j].
 * mwd_hmm_gauss_emission fills emis (n_slots) and the log mixture responsibilities resp
 * (n_slots x M); lnorm [dev] Vt x M is scratch.                                              */
int mwd_hmm_gauss_emission(const mwd_hmm_problem* p, const void* emb, int emb_is_f64, int emb_dim,
                           int n_mix, const double* lprior, const double* means, const double* var,
                           double* lnorm, double* emis, double* resp, void* stream);
/* Posterior-weighted sufficient statistics per (word, mixture): stats[w][m] = [sum wgt,
 * sum wgt*x (D), sum wgt*x^2 (D)], wgt = exp(post + resp); slots are visited through a postings
 * index sorted by word (word_idx, word_off[Vt+1]) in fixed order (smt/...:339-375).          */
int mwd_hmm_gauss_stats(const mwd_hmm_problem* p, const void* emb, int emb_is_f64, int emb_dim,
                        int n_mix, const double* resp, const int64_t* word_idx, const int64_t* word_off,
                        double* stats, void* stream);
/* means = sum wgt*x / sum wgt; mixture priors renormalised when n_mix > 1; variances only when
 * update_var != 0 (the reference's fixedVariance > 0 keeps them).                            */
int mwd_hmm_gauss_update(int n_tgt_types, int n_mix, int emb_dim, const double* stats, int update_var,
                         double* lprior, double* means, double* var, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MWD_B200_H */
