/*
 * mds_adf.h — C ABI of the B200-native angular-distribution-function path,
 * exported by the same shared library as mds_rdf.h (mdsuite_b200/_cuda/libmds_rdf.so).
 *
 * The reference (zincware/MDSuite 0.2.0) has no FFI; the seam this ABI replaces is the part of
 *   mdsuite/calculators/angular_distribution_function.py:374-405   (_build_histograms)
 * that runs per batch of configurations BEFORE np.histogram normalises:
 *   get_neighbour_list (utils/neighbour_list.py:36-101)  -> minimum-image r_ij
 *   get_triplets       (utils/neighbour_list.py:104-177) -> float16 cutoff, all (i, j, k) triples
 *   _compute_angles    (:341-372)                        -> species combination of a triple
 *   get_angles         (utils/linalg.py:53-81)           -> theta = acos(u_ij . u_ik),
 *                                                           weight = 1 / (|r_ij| |r_ik|)^norm_power
 *   np.histogram(theta, bins, range=[0, 3.15], weights)  -> sum of weights per bin
 * Output: for every configuration and every species combination (A, B, C), A <= B <= C in
 * itertools.combinations_with_replacement order, the sum of the weights per bin as float64
 * (integers when norm_power = 0).  Density normalisation per batch, the float32 accumulation over
 * batches and the x axis stay with the caller (the calculator), as in the reference.
 *
 * Inputs are what that code consumes: positions float32 [n_frames][n_atoms][3] with the species
 * concatenated in `args.species` order (_format_data :459-487), the atoms per species
 * (_prepare_data_structure :262-301), box_array, cutoff, number_of_bins, norm_power.
 *
 * Conventions as in mds_rdf.h: plain C, MDS_OK (0) or a negative status, thread-local error text,
 * the library owns its device memory, one handle per (thread, GPU), no CPU fallback.
 */
#ifndef MDS_ADF_H
#define MDS_ADF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDS_ADF_ABI_VERSION 1

#if defined(MDS_BUILD) && defined(__GNUC__)
#define MDS_ADF_API __attribute__((visibility("default")))
#else
#define MDS_ADF_API
#endif

#define MDS_ADF_OK 0
#define MDS_ADF_ERR_INVALID (-1)
#define MDS_ADF_ERR_CUDA (-2)
#define MDS_ADF_ERR_NOMEM (-3)
#define MDS_ADF_ERR_UNSUPPORTED (-4)

#define MDS_ADF_MAX_NEIGHBOURS 1024 /* neighbours of one atom that can take part in a triple */
#define MDS_ADF_RANGE_HI 3.15      /* upper edge of the angle axis, "a chemist's pi" (:215) */

typedef struct mds_adf_config {
  uint32_t struct_size;   /* sizeof(mds_adf_config) */
  int32_t n_species;      /* S >= 1 */
  const int64_t* counts;  /* [S] atoms per species, storage order */
  int32_t nbins;          /* number_of_bins >= 1 */
  float cutoff;           /* r_cut > 0; compared in float16 like the reference */
  float box[3];           /* float32(box_array) > 0, orthorhombic */
  int32_t norm_power;     /* 0..16 */
  int32_t device;         /* CUDA device ordinal */
  int32_t max_frames_in_flight; /* configurations per internal batch; <= 0: library default */
} mds_adf_config;

typedef struct mds_adf_stats {
  uint32_t struct_size;
  int32_t n_combinations;   /* S (S + 1) (S + 2) / 6 */
  int64_t frames;           /* configurations processed since create */
  double triples;           /* ordered (i, j, k) triples binned since create */
  double pair_tests;        /* (i, j) distance tests of the neighbour search since create, counted
                               on the device */
  int64_t kernel_launches;
  float kernel_ms;          /* device time of the kernels of the LAST compute call (CUDA events) */
  float compute_ms;         /* device time of the whole LAST compute call incl. copies */
  int32_t max_neighbours;   /* largest neighbour list seen since create */
  int32_t sm_count;
  int32_t cells[3];         /* cells per dimension of the neighbour search (0: every atom is tested
                               against every atom — small boxes, few atoms) */
  int32_t reserved;
  int64_t frames_far;       /* cell search: configurations with a coordinate beyond 8 box lengths,
                               which were searched atom against atom instead */
} mds_adf_stats;

typedef struct mds_adf mds_adf_t;

MDS_ADF_API int mds_adf_create(mds_adf_t** out, const mds_adf_config* cfg);
MDS_ADF_API void mds_adf_destroy(mds_adf_t* h);

/* Weighted angle histograms of n_frames configurations in HOST memory, xyz = [n_frames][n_atoms][3]
 * float32.  Writes out[n_frames][n_combinations][nbins] float64 and returns when it is complete.
 * MDS_ADF_ERR_UNSUPPORTED if an atom has more than MDS_ADF_MAX_NEIGHBOURS neighbours. */
MDS_ADF_API int mds_adf_compute(mds_adf_t* h, const float* xyz, int64_t n_frames, double* out);

/* Same with positions and output in DEVICE memory on the handle's device; ordered after `stream`
 * (cudaStream_t as void*, NULL = default stream), returns when the result is complete. */
MDS_ADF_API int mds_adf_compute_device(mds_adf_t* h, const float* d_xyz, int64_t n_frames,
                                       double* d_out, void* stream);

MDS_ADF_API int mds_adf_stats_get(mds_adf_t* h, mds_adf_stats* out);

/* The cosine thresholds the kernel bins with (host only, no CUDA call): cos_edges[k], k = 1 ..
 * nbins - 1, is the largest float32 c whose angle float32(acos(c)) np.histogram puts in bin >= k
 * (-2 if there is none: bins beyond pi); cos_edges[0] = +inf, cos_edges[nbins] = -2.  A clipped
 * cosine c falls in bin #{k >= 1 : c <= cos_edges[k]}. */
MDS_ADF_API int mds_adf_build_threshold_table(int nbins, float* cos_edges /* [nbins + 1] */);

/* The squared-distance window of the neighbour test (host only): j is a neighbour of i iff
 * *s_lo <= |r_ij|^2 < *s_hi, which is float16(|r_ij|) != 0 and float16(|r_ij|) < float16(cutoff)
 * with |r_ij| = sqrt_rn(|r_ij|^2). */
MDS_ADF_API int mds_adf_neighbour_window(float cutoff, float* s_lo, float* s_hi);

MDS_ADF_API const char* mds_adf_last_error(void);
MDS_ADF_API int mds_adf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDS_ADF_H */
