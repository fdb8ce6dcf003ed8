/*
 * mds_rdf.h — C ABI of the B200-native radial-distribution-function hot path.
 *
 * The reference (zincware/MDSuite 0.2.0) is pure Python over TensorFlow and has NO FFI; the
 * seam this ABI replaces is the body of the batch loop of
 *   mdsuite/calculators/radial_distribution_function.py:846-885   (run_calculator, HOT LOOP 1)
 * i.e. run_minibatch_loop (:422-468) -> get_partial_triu_indices (utils/linalg.py:102-122)
 * -> get_dij (:647-689) -> apply_minimum_image (utils/linalg.py:84-99)
 * -> compute_species_values (:470-524) -> bin_minibatch (:616-645)
 * -> apply_system_cutoff (utils/linalg.py:125-136) -> tf.histogram_fixed_width (:642-644)
 * -> combine_dictionaries (:601-614) and the per-key accumulation `self.rdf[key] += rdf[key]`
 * (:884-885).
 *
 * Inputs are what that loop body consumes:
 *   positions_tensor  float32, species concatenated in `args.species` order
 *                     (_format_data :535-563)          -> `xyz`, layout MDS_LAYOUT_*
 *   particles_list    (:691-717)                       -> mds_rdf_config.counts
 *   box_array cast to float32 (:450)                   -> mds_rdf_config.box
 *   cutoff cast to float32 (:522), number_of_bins (:521) -> mds_rdf_config.cutoff / nbins
 * Output is what it produces: one integer histogram per species pair in
 * itertools.combinations_with_replacement order (:272-275, :502), here uint64 instead of the
 * reference's wrapping int32 (quirk Q5).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns MDS_OK (0) or a negative
 *     mds_status; mds_last_error() returns a thread-local message for the last failure.
 *   - the library owns all device memory, streams and events of a handle; the caller owns
 *     host buffers and must keep a submitted host buffer alive until the next
 *     mds_rdf_counts(sync=1) / mds_rdf_synchronize().
 *   - one handle per (thread, GPU); a handle is not thread-safe, different handles are
 *     independent.
 *   - there is no CPU fallback: without a usable CUDA device mds_rdf_create fails.
 */
#ifndef MDS_RDF_H
#define MDS_RDF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDS_RDF_ABI_VERSION 7

#if defined(MDS_BUILD) && defined(__GNUC__)
#define MDS_API __attribute__((visibility("default")))
#else
#define MDS_API
#endif

typedef enum mds_status {
  MDS_OK = 0,
  MDS_ERR_INVALID = -1,     /* bad argument */
  MDS_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  MDS_ERR_NOMEM = -3,       /* allocation failed */
  MDS_ERR_UNSUPPORTED = -4  /* valid request this build cannot serve */
} mds_status;

/* mds_rdf_config.flags */
#define MDS_RDF_PARITY_SKIP_FIRST 0x1u /* reproduce quirk Q1: the atom at each species' offset is
                                          never a row nor a column (strict '>' at :637-638) */
#define MDS_RDF_PATH_ALLPAIRS 0x2u     /* force the tiled all-pairs kernel */
#define MDS_RDF_PATH_CELLLIST 0x4u     /* force the cell-list kernel (needs cutoff <= box/3.003;
                                          MDS_ERR_UNSUPPORTED otherwise) */
/* neither PATH bit set: the library picks per handle from n_atoms, box and cutoff */

/* layouts of the xyz argument (float32) */
#define MDS_LAYOUT_FRAME_MAJOR_XYZ 0 /* [n_frames][n_atoms][3]  — what the frame batcher streams */
#define MDS_LAYOUT_ATOM_MAJOR_XYZ 1  /* [n_atoms][n_frames][3]  — the reference's positions_tensor
                                        (N, C, 3) as loaded from its HDF5 store */

typedef struct mds_rdf_config {
  uint32_t struct_size;        /* sizeof(mds_rdf_config), for ABI evolution */
  int32_t n_species;           /* S >= 1 */
  const int64_t* counts;       /* [S] atoms per species (particles_list) */
  int32_t nbins;               /* B >= 1 */
  float cutoff;                /* float32(cutoff) > 0 */
  float box[3];                /* float32(box_array) > 0, orthorhombic */
  uint32_t flags;              /* MDS_RDF_* */
  int32_t device;              /* CUDA device ordinal */
  int32_t max_frames_in_flight; /* frames per internal batch; <= 0: library default */
} mds_rdf_config;

typedef struct mds_rdf_stats {
  uint32_t struct_size;
  int32_t path;                /* MDS_RDF_PATH_ALLPAIRS or MDS_RDF_PATH_CELLLIST actually used */
  int64_t frames;              /* frames accumulated since create/reset */
  double pairs_evaluated;      /* (i, j, frame) triples distance-tested by this handle, counted on the
                                  device: reference pairs on the all-pairs path, candidate pairs of
                                  adjacent cells on the cell path (this shard's share if sharded) */
  double pairs_allpairs_equiv; /* frames * sum over species pairs of the reference's pair count
                                  (divided by the shard count if sharded) */
  double pairs_binned;         /* triples with d < cutoff (= sum of all histogram bins) */
  int64_t frames_general_minimage; /* frames that needed the div/rint minimum image (not wrapped) */
  int64_t kernel_launches;     /* kernels of this library launched since create/reset */
  float kernel_ms;             /* device time of the pair kernels of the LAST submit (CUDA events) */
  float submit_ms;             /* device time of the whole LAST submit incl. copies and packing */
  int32_t n_pairs;             /* S(S+1)/2 */
  int32_t nbins;
  int32_t sm_count;
  int32_t variant;             /* kernel variant picked for this handle (tuning aid) */
  float pair_kernel_ms;        /* sum of pair-kernel launch durations since the previous stats_get */
  int32_t pair_kernel_launches; /* pair-kernel launches since the previous stats_get */
  int32_t dense;               /* 1: branch-free binning (most pairs inside the cutoff) */
  int32_t frames_far;          /* cell path: frames handed to the all-pairs kernel — a coordinate
                                  beyond 8 box lengths, or a cell too crowded to stage */
  int32_t cells[3];            /* cell path: cells per dimension (0 on the all-pairs path) */
  int32_t item_shards;         /* shard count set with mds_rdf_set_item_shard (1: none) */
  int32_t cell_stencil[3];     /* cell path: stencil half-widths — the cell kernel tests atoms whose
                                  cells are at most this many cells apart (1 or 2 per dimension) */
  int32_t cell_build;          /* cell path, last batch: 1 = lists built by the one-kernel build (a CTA
                                  per frame, counters in shared memory), 0 = pack / scan / scatter */
  double pairs_tested;         /* distances actually computed: = pairs_evaluated on the all-pairs
                                  path; on the cell path the kernel works on 64-atom chunks and tests
                                  a superset of the candidate pairs (cell kernel launches only) */
} mds_rdf_stats;

typedef struct mds_rdf mds_rdf_t;

/* Create a handle: allocates device buffers and streams, builds the squared-distance threshold
 * table that replaces sqrt -> (d < cutoff) -> histogram_fixed_width bit-exactly. */
MDS_API int mds_rdf_create(mds_rdf_t** out, const mds_rdf_config* cfg);
MDS_API void mds_rdf_destroy(mds_rdf_t* h);

/* Accumulate the histograms of n_frames configurations held in HOST memory (replaces one or more
 * iterations of the batch loop :846-885).  Asynchronous.
 *   host_is_pinned != 0  promises page-locked memory: the frames are copied straight out of the
 *     caller's buffer on a side stream in max_frames_in_flight chunks, double buffered against the
 *     kernels; the buffer must stay untouched until mds_rdf_wait_host_copies / a synchronising call.
 *   host_is_pinned == 0  (pageable memory, e.g. a NumPy array per batch): the calling thread copies
 *     the frames into the library's own ring of pinned buffers (transposing the atom-major layout
 *     to frame-major on the way); a buffer travels to the device when it is full or when results
 *     are asked for.  The caller's buffer may be reused as soon as the call returns, and many
 *     small calls (one configuration each, as the reference batches at default memory settings)
 *     cost one memcpy each instead of one transfer + four kernel launches each. */
MDS_API int mds_rdf_submit(mds_rdf_t* h, const float* xyz, int64_t n_frames, int layout,
                   int host_is_pinned);

/* Same for configurations already in DEVICE memory on the handle's device.  `stream` is the
 * cudaStream_t (as void*) on which the data was produced, or NULL for the default stream; the
 * library orders its work after it and never blocks the host. */
MDS_API int mds_rdf_submit_device(mds_rdf_t* h, const float* d_xyz, int64_t n_frames, int layout,
                          void* stream);

/* Copy the accumulated counts [n_pairs][nbins] (uint64) to host memory.  sync != 0 waits for all
 * submitted work first (and must be used before the host buffers of mds_rdf_submit are reused);
 * sync == 0 enqueues the copy and returns (out must be pinned; call mds_rdf_synchronize). */
MDS_API int mds_rdf_counts(mds_rdf_t* h, uint64_t* out, int sync);

/* Device pointer to the accumulated counts [n_pairs][nbins] uint64 and the stream they are
 * produced on, for a collective over per-GPU handles (the host layer allreduces this buffer in
 * place with NCCL, one ncclAllReduce(sum, uint64) — SURVEY.md §8e). */
MDS_API int mds_rdf_device_counts(mds_rdf_t* h, void** d_counts, void** stream);

/* The one collective of the path: sum the accumulated counts over the ranks of
 * `comm` (an ncclComm_t of NCCL 2.x; every rank calls this with its own handle of the same
 * configuration), in place, as ONE ncclAllReduce(sum, ncclUint64, n_pairs * nbins) on the library's
 * stream — ordered after everything submitted so far and before anything submitted or reset
 * afterwards, without blocking the host.  Afterwards mds_rdf_counts returns the global histogram
 * on every rank.  NCCL is resolved at run time (no link-time dependency): MDS_ERR_UNSUPPORTED if
 * ncclAllReduce is neither loaded in the process nor found as libnccl.so.2. */
struct ncclComm;
MDS_API int mds_rdf_allreduce(mds_rdf_t* h, struct ncclComm* comm);

/* Make `stream` (a cudaStream_t as void*, NULL = default stream) wait for everything submitted so
 * far, without blocking the host — so that events recorded on the caller's stream bracket the
 * library's work. */
MDS_API int mds_rdf_join(mds_rdf_t* h, void* stream);

/* The reverse: make the library's stream wait for work already enqueued on `stream` (e.g. a
 * collective that reads the device counts) before anything submitted or reset afterwards. */
MDS_API int mds_rdf_wait(mds_rdf_t* h, void* stream);

/* Block the host until every host->device copy issued by mds_rdf_submit so far has completed,
 * i.e. until the submitted host buffers may be overwritten (the kernels may still be running). */
MDS_API int mds_rdf_wait_host_copies(mds_rdf_t* h);

/* Finer than the above, for a feeder with a ring of pinned buffers: mds_rdf_host_ticket returns
 * the number of page-locked host submits made so far (the ticket of the latest one; 0: none),
 * mds_rdf_wait_host_ticket blocks until the copies of submit number `ticket` have left the host
 * buffer — later submits may still be in flight, so refilling buffer k overlaps the copy of
 * buffer k + 1. */
MDS_API int64_t mds_rdf_host_ticket(mds_rdf_t* h);
MDS_API int mds_rdf_wait_host_ticket(mds_rdf_t* h, int64_t ticket);

/* How mds_rdf_submit cuts a call of n_frames host frames into batches (each batch: one
 * host->device copy on the copy stream, overlapped with the kernels of the batch before it; the
 * reference's counterpart is the np.array_split of calculators/radial_distribution_function.py:594
 * into memory-sized batches).  batch_frames: the handle's full batch; compute_idle != 0: nothing is
 * running on the handle when the call arrives, so the first copy cannot be hidden and the batches
 * start at batch_frames / 16 and grow by x 1.5 (cell_path != 0) or x 4 until a full batch is
 * reached; the rest of the call is split into equal batches of at most batch_frames.  Pure host
 * arithmetic (no CUDA call, works without a GPU).  Writes up to `capacity` batch sizes to `out`
 * and returns how many batches the plan has (negative: invalid argument). */
MDS_API int64_t mds_rdf_plan_host_batches(int64_t n_frames, int32_t batch_frames, int compute_idle,
                                          int cell_path, int64_t* out, int64_t capacity);

/* Item sharding — strong scaling WITHIN frames, for runs with fewer configurations than GPUs
 * (SURVEY.md §8e "fallback for F < G"): every rank submits the same frames, the handle of rank
 * `index` of `count` processes every count-th unit of the (work item, frame) enumeration, and the
 * allreduce over the ranks' counts gives the full histogram.  Affects submits made afterwards;
 * (0, 1) restores the default.  Statistics report this shard's share. */
MDS_API int mds_rdf_set_item_shard(mds_rdf_t* h, int32_t index, int32_t count);

MDS_API int mds_rdf_synchronize(mds_rdf_t* h);
MDS_API int mds_rdf_reset(mds_rdf_t* h); /* zero the accumulated counts and statistics */
MDS_API int mds_rdf_stats_get(mds_rdf_t* h, mds_rdf_stats* out);

/* The squared-distance thresholds the kernels bin with: edges[k], k = 0..nbins, is the smallest
 * float32 s = |r|^2 that the reference's sqrt + histogram would put in bin >= k (edges[0] = 0,
 * edges[nbins] = *s_cut); *s_cut is the smallest float32 s with sqrt(s) >= cutoff, i.e. a pair
 * is inside the cutoff iff s < *s_cut.  For tests and audits. */
MDS_API int mds_rdf_threshold_table(mds_rdf_t* h, float* edges /* [nbins + 1] */, float* s_cut);

/* Handle-free, host-only version of the above (no CUDA call): fills edges[0..nbins] and *s_cut
 * for float32(cutoff) and nbins.  edges[nbins] == *s_cut. */
MDS_API int mds_rdf_build_threshold_table(float cutoff, int nbins, float* edges, float* s_cut);

/* Micro-benchmark behind the shared-atomic roofline: conflict-free and
 * random-address shared-memory atomicAdd lanes per clock per SM on `device`. */
MDS_API int mds_bench_shared_atomics(int device, int nbins, int iters, double* lanes_per_clk_per_sm_spread,
                             double* lanes_per_clk_per_sm_random, double* sm_clock_mhz);

MDS_API const char* mds_last_error(void);
MDS_API int mds_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDS_RDF_H */
