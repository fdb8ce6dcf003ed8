/*
 * mds_io.h — C ABI of the native trajectory ingestion (host code, no CUDA): LAMMPS text dumps and
 * extended XYZ files.
 *
 * Replaces the pure-Python text parsing of the reference on the way INTO the trajectory store
 * (SURVEY.md §8f rank 2):
 *   mdsuite/file_io/lammps_trajectory_files.py:103-146   header analysis, n_configs from the line
 *                                                        count, species of the first configuration
 *   mdsuite/file_io/tabular_text_files.py:160-220        per configuration: skip 9 header lines,
 *                                                        split n_particles lines, sort by the id
 *                                                        column (utils/meta_functions.py:519-527),
 *                                                        slice property columns
 *   mdsuite/file_io/extxyz_files.py:93-121               extended XYZ: atom count + key=value line
 *                                                        per configuration, atoms in file order
 * The reference turns every token into float64 (NumPy's correctly rounded string conversion) and
 * stores float32 (h5py dataset dtype, database/simulation_database.py:491-497); this library
 * returns float32((double) token) — the same two roundings.
 *
 * Conventions as in mds_rdf.h: plain C, 0 / negative status, thread-local error string.
 */
#ifndef MDS_IO_H
#define MDS_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDS_IO_ABI_VERSION 3

#if defined(MDS_BUILD) && defined(__GNUC__)
#define MDS_IO_API __attribute__((visibility("default")))
#else
#define MDS_IO_API
#endif

#define MDS_IO_OK 0
#define MDS_IO_ERR_INVALID (-1)
#define MDS_IO_ERR_IO (-2)
#define MDS_IO_ERR_FORMAT (-3)
#define MDS_IO_ERR_NOMEM (-4)

typedef struct mds_dump mds_dump_t;

typedef struct mds_dump_info {
  uint32_t struct_size;
  int32_t n_columns;        /* per-atom columns named on the "ITEM: ATOMS" line */
  int64_t n_atoms;          /* header line 4 */
  int64_t n_configurations; /* total lines / (n_atoms + 9), must divide */
  int64_t file_bytes;
  int64_t timestep0;        /* header line 2 of the first configuration */
  int64_t timestep1;        /* ... of the second one (== timestep0 if there is only one) */
  double box_lo[3];         /* header lines 6-8 */
  double box_hi[3];
  int32_t id_column;        /* index of "id", -1 if absent */
  int32_t species_column;   /* index of "element", else "type", else -1 */
} mds_dump_info;

/* mmap the dump, read the first header and index the configurations with n_threads threads
 * (<= 0: hardware concurrency). */
MDS_IO_API int mds_dump_open(mds_dump_t** out, const char* path, int n_threads);
/* Same for an extended XYZ file (2 header lines per configuration: the atom count and the
 * key=value line).  The handle serves every mds_dump_* call below; in its info n_columns is the
 * token count of the first atom line (columns are named "c0", "c1", ...), id_column and
 * species_column are -1 and the timestep / box fields are 0: Lattice, Properties and time live on
 * the key=value line, which the caller reads with mds_dump_header_line(h, frame, 1, ...). */
MDS_IO_API int mds_xyz_open(mds_dump_t** out, const char* path, int n_threads);
MDS_IO_API void mds_dump_close(mds_dump_t* h);
MDS_IO_API int mds_dump_info_get(mds_dump_t* h, mds_dump_info* out);
/* name of per-atom column idx into buf (NUL terminated, truncated to buf_len) */
MDS_IO_API int mds_dump_column_name(mds_dump_t* h, int idx, char* buf, int buf_len);
/* Header line `line` (0-based; 9 per configuration in a LAMMPS dump, 2 in an extended XYZ file) of
 * configuration `frame`, without its line end, NUL terminated and truncated to buf_len - 1 bytes.
 * Returns the full length of the line (>= 0), or a negative status. */
MDS_IO_API int64_t mds_dump_header_line(mds_dump_t* h, int64_t frame, int line, char* buf,
                                        int64_t buf_len);
/* Labels of column `column` for the first configuration, atoms in ascending id order (file
 * order if sort_by_id == 0): n_atoms NUL-terminated strings at stride label_stride bytes. */
MDS_IO_API int mds_dump_first_frame_labels(mds_dump_t* h, int column, int sort_by_id, char* buf,
                                           int label_stride);
/* out[frame][atom][c] = float32((double) token of column columns[c]) for configurations
 * [frame0, frame0 + n_frames), atoms in ascending id order (file order if sort_by_id == 0). */
MDS_IO_API int mds_dump_read_columns(mds_dump_t* h, int64_t frame0, int64_t n_frames,
                                     const int32_t* columns, int32_t n_cols, int sort_by_id,
                                     float* out);
/* Same, with the atom of rank k (k-th smallest id, or k-th line if sort_by_id == 0) written to
 * row dest[k] (a permutation of 0..n_atoms-1; NULL = identity) — lets the caller receive the
 * species-contiguous order of the trajectory store without a second gather. */
MDS_IO_API int mds_dump_read_columns_mapped(mds_dump_t* h, int64_t frame0, int64_t n_frames,
                                            const int32_t* columns, int32_t n_cols, int sort_by_id,
                                            const int32_t* dest, float* out);
/* The string -> float32 conversion on its own (for parity tests): n NUL-terminated tokens at
 * stride `stride` bytes -> out[n]. */
MDS_IO_API int mds_parse_float32(const char* tokens, int64_t n, int32_t stride, float* out);

/* ---- HDF5 chunk decoding: the feed out of a reference project's database.hdf5 -------------------
 * The reference stores <species>/Positions as (n_atoms, n_configurations, 3) float32, chunked and
 * gzip-compressed (mdsuite/database/simulation_database.py:449-499) and reads it with h5py fancy
 * indexing (:594-639): libhdf5 inflates chunk after chunk on one thread.  The host language walks
 * the file's metadata (mdsuite_b200/database/hdf5_reader.py) and hands the chunks of a request to
 * mds_h5_read_chunks, which decodes them on all cores and writes the requested box with the
 * caller's strides — e.g. frame-major for the GPU feed. */
typedef struct mds_h5_chunk {
  int64_t address;        /* file offset of the stored (filtered) chunk */
  int64_t nbytes;         /* its stored size */
  int64_t offsets[4];     /* dataset coordinates of the chunk's first element (rank entries used) */
  uint32_t filter_mask;   /* bit i set: filter i of the pipeline was skipped for this chunk */
  uint32_t reserved;
} mds_h5_chunk;

typedef struct mds_h5_layout {
  uint32_t struct_size;   /* sizeof(mds_h5_layout) */
  int32_t rank;           /* 1..4 */
  int32_t elem_size;      /* bytes per element */
  int32_t n_filters;      /* filter pipeline in write order: 1 deflate, 2 shuffle, 3 fletcher32 */
  int32_t filter_ids[8];
  int32_t filter_cd0[8];  /* first client-data value of each filter (shuffle: element size) */
  int32_t verify_checksums; /* != 0: check the adler32 of every deflate stream */
  int32_t reserved;
  int64_t chunk_dims[4];
  int64_t box_lo[4];      /* requested box [lo, hi) in dataset coordinates */
  int64_t box_hi[4];
  int64_t out_strides[4]; /* element strides of the output: index = sum (i_d - lo_d) * stride_d */
} mds_h5_layout;

/* Decode `n_chunks` chunks read from file descriptor `fd` (pread; the descriptor's offset is not
 * used) with n_threads threads (<= 0: hardware concurrency) and copy their intersection with the
 * box into `out`.  Elements of the box that no chunk covers are left untouched (the caller
 * pre-fills the HDF5 fill value). */
MDS_IO_API int mds_h5_read_chunks(int fd, const mds_h5_layout* layout, const mds_h5_chunk* chunks,
                                  int64_t n_chunks, void* out, int n_threads);
/* The library's own inflate on one zlib stream (RFC 1950 / 1951), for tests and audits. */
MDS_IO_API int mds_inflate(const void* src, int64_t src_len, void* dst, int64_t dst_capacity,
                           int64_t* dst_len, int verify_checksum);

MDS_IO_API const char* mds_io_last_error(void);
MDS_IO_API int mds_io_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDS_IO_H */
