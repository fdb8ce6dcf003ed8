/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement ("oracle") of the MDSuite RDF hot path.
 *
 *   PARITY: control flow, species masks (quirk Q1) and accumulation are pinned against the
 *   reference's own run_calculator executed over a NumPy stand-in for its TensorFlow leaf ops
 *   (tests/golden/reference_golden.json, tests/test_reference_golden.py); the float32 leaf
 *   arithmetic lives in TensorFlow (not vendored, not importable here) and is a restatement.
 *   See oracle/rdf_oracle.py for the op-by-op NumPy twin (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library built from this file.  The product never links it.
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off -fopenmp, see oracle/Makefile)
 *
 * Reference lines followed (relative to /root/reference/mdsuite/):
 *   utils/linalg.py:102-122                       strict upper triangle  col > row
 *   calculators/radial_distribution_function.py:667-676   r = pos[col] - pos[row]
 *   utils/linalg.py:99                            r - rint(r / box) * box   (4 rounded f32 ops)
 *   calculators/radial_distribution_function.py:686       sqrt((x*x + y*y) + z*z)
 *   calculators/radial_distribution_function.py:637-638   species masks with strict '>' (quirk Q1)
 *   utils/linalg.py:134                           d < float32(cutoff)
 *   calculators/radial_distribution_function.py:642-644   tf.histogram_fixed_width:
 *        step = double(float(hi) - float(lo)) / double(B);
 *        bin  = (int32) min(double(max(d, lo) - lo) / step, double(B - 1))
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_LAYOUT_ATOM_MAJOR 0  /* (N, C, 3): the reference's positions_tensor */
#define ORACLE_LAYOUT_FRAME_MAJOR 1 /* (C, N, 3) */

static inline float minimum_image_1d(float r, float box) {
  float q = r / box;    /* RealDiv */
  float n = rintf(q);   /* Rint, half to even under the default rounding mode */
  float nb = n * box;   /* Mul */
  return r - nb;        /* Sub */
}

static inline int64_t pos_index(int layout, int64_t n_atoms, int64_t n_frames, int64_t atom,
                                int64_t frame) {
  return layout == ORACLE_LAYOUT_ATOM_MAJOR ? (atom * n_frames + frame) * 3
                                            : (frame * n_atoms + atom) * 3;
}

/* Squared minimum-image distance exactly as the reference rounds it. */
float rdf_oracle_s(const float* pi, const float* pj, const float* box) {
  float rx = minimum_image_1d(pj[0] - pi[0], box[0]);
  float ry = minimum_image_1d(pj[1] - pi[1], box[1]);
  float rz = minimum_image_1d(pj[2] - pi[2], box[2]);
  float xx = rx * rx, yy = ry * ry, zz = rz * rz;
  float s = xx + yy;
  s = s + zz;
  return s;
}

/*
 * counts: (n_pairs, nbins) int64, zeroed here.  Species pairs are ordered like
 * itertools.combinations_with_replacement(range(S), 2).  skip_first != 0 reproduces quirk Q1.
 * Returns the number of (row, col, frame) triples evaluated, or -1 on bad arguments.
 */
int64_t rdf_oracle_counts(const float* pos, int64_t n_atoms, int64_t n_frames, int layout,
                          const int64_t* particles, int n_species, const float* box,
                          float cutoff, int nbins, int skip_first, int64_t* counts,
                          int n_threads) {
  if (!pos || !particles || !box || !counts || nbins < 1 || n_species < 1) return -1;
  int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_species + 1));
  offs[0] = 0;
  for (int s = 0; s < n_species; ++s) offs[s + 1] = offs[s] + particles[s];
  if (offs[n_species] != n_atoms) { free(offs); return -1; }
  const int n_pairs = n_species * (n_species + 1) / 2;
  memset(counts, 0, sizeof(int64_t) * (size_t)n_pairs * (size_t)nbins);
  const float lo = 0.0f, hi = cutoff;
  const double step = (double)(hi - lo) / (double)nbins;
  const double last = (double)(nbins - 1);
  const int first = skip_first ? 1 : 0;
  int64_t evaluated = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : evaluated)
  {
    int64_t* local = (int64_t*)calloc((size_t)n_pairs * (size_t)nbins, sizeof(int64_t));
    int pair = 0;
    for (int a = 0; a < n_species; ++a) {
      for (int b = a; b < n_species; ++b, ++pair) {
        int64_t* h = local + (int64_t)pair * nbins;
        const int64_t r0 = offs[a] + first, r1 = offs[a + 1];
        const int64_t c0 = offs[b] + first, c1 = offs[b + 1];
#pragma omp for schedule(dynamic, 8) collapse(2) nowait
        for (int64_t f = 0; f < n_frames; ++f) {
          for (int64_t row = r0; row < r1; ++row) {
            const float* pi = pos + pos_index(layout, n_atoms, n_frames, row, f);
            int64_t cs = (a == b) ? (row + 1 > c0 ? row + 1 : c0) : c0;
            for (int64_t col = cs; col < c1; ++col) {
              const float* pj = pos + pos_index(layout, n_atoms, n_frames, col, f);
              float d = sqrtf(rdf_oracle_s(pi, pj, box));
              ++evaluated;
              if (d < cutoff) {
                float v = (d > lo ? d : lo) - lo;
                double q = (double)v / step;
                if (q > last) q = last;
                h[(int32_t)q] += 1;
              }
            }
          }
        }
      }
    }
#pragma omp critical
    {
      for (int64_t i = 0; i < (int64_t)n_pairs * nbins; ++i) counts[i] += local[i];
    }
    free(local);
  }
  free(offs);
  return evaluated;
}

int rdf_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
