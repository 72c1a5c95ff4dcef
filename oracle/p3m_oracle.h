/* oracle/p3m_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's P3M / PM force step (the hot path named by
 * BASELINE.json:north_star).  Every function cites the reference
 * file:line it follows.  Two symbol sets are exported from one source: *_f32 computes in float
 * exactly like the reference (which is fp32-only), *_f64 is the same algorithm in double on the
 * same fp32 inputs (the 1e-6 bar of north_star needs it).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library, and
 * only as the checker.  The product (particlesimulation_b200/, include/p3m_b200.h) never does.
 *
 * Parity pin: the reference's own tests hold no golden vector for this path, so
 * the restatement is pinned against the UNMODIFIED reference compiled here (oracle/_ref, built by
 * oracle/Makefile from /root/reference) -- tests/test_oracle_vs_ref.py -- and against the fixtures
 * that build generated, committed under tests/golden/ (tests/golden/make_golden.py).
 */
#ifndef P3M_ORACLE_H
#define P3M_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as RefParams in oracle/ref_driver.cpp; all physical inputs are fp32 like the
 * reference's constructor arguments (include/pmMethod.h:14-26, include/p3mMethod.h:13-27). */
typedef struct OrcParams {
  int n;
  int nx, ny, nz;
  float box[3];
  float H, DT, G;
  int is;    /* 0 NGP 1 CIC 2 TSC */
  int fds;   /* 0 TWO_POINT 1 FOUR_POINT */
  int gfunc; /* 0 DISCRETE_LAPLACIAN 1 S1_OPTIMAL 2 S2_OPTIMAL 3 POOR_MAN */
  float particleDiameter;
  float cutoffRadius;
  float softening;
  int cloudShape; /* 0 S1 1 S2 */
  int useTable;
  int ySort;
  int extKind; /* 0 none, 1 sphRadDecrField */
  float extCenter[3];
  float extR, extM;
  /* 1: G := 0 at the modes where D^ vanishes identically (every k_i in {0, N_i/2}); there
   * GreenOptimal is 0/0 rounding noise (+0.153 in fp32, -1.1e8 in fp64 at (16,16,16) of
   * a 32^3 mesh) and the mode carries no force.  0 = literal reference behaviour (fp32 pin). */
  int greenZeroDegenerate;
} OrcParams;

#define ORC_DECL(R, S)                                                                             \
  /* unitConversions.h:8-50, unitConversions.cpp:23-71: original -> code units */                  \
  void orc_to_code_units_##S(const OrcParams* p, const float* pos, const float* vel,               \
                             const float* mass, R* pos_c, R* vel_c, R* mass_c);                    \
  /* pmMethod.cpp:164-185 + greensFunctions.cpp:122-220: influence function, real part, M values   \
   * x fastest */                                                                                  \
  void orc_green_##S(const OrcParams* p, R* green);                                                \
  /* pmMethod.cpp:200-277 + grid.cpp:28-36: mass assignment (code units) */                        \
  void orc_deposit_##S(const OrcParams* p, const R* pos_c, const R* mass_c, R* density);           \
  /* grid.cpp:50-56, pmMethod.cpp:340-350: C2C FFT, multiply by G, inverse C2C / M, real part */   \
  void orc_poisson_##S(const OrcParams* p, const R* density, const R* green, R* potential);        \
  /* pmMethod.cpp:352-382: periodic 2-/4-point differences; field = 3M interleaved xyz */          \
  void orc_field_##S(const OrcParams* p, const R* potential, R* field);                            \
  /* pmMethod.cpp:279-338,384-390: interpolate field + external field; acc = 3N */                 \
  void orc_gather_##S(const OrcParams* p, const R* pos_c, const R* field, R* acc);                 \
  /* p3mMethod.cpp:275-294: 500-entry short-range force table */                                   \
  void orc_sr_table_##S(const OrcParams* p, R* table500);                                          \
  /* chainingMesh.cpp:6-29,79-84: dims[3] and flat cell of every particle (-1 outside) */          \
  void orc_chaining_cells_##S(const OrcParams* p, const R* pos_c, int* dims, int* cell);           \
  /* chainingMesh.cpp:20-58: particle ids cell by cell in the reference's list order */            \
  void orc_chaining_order_##S(const OrcParams* p, const R* pos_c, int* order);                     \
  /* p3mMethod.cpp:168-192,220-273,296-322: total short-range force per particle (own + 13 slots)  \
   */                                                                                              \
  void orc_sr_forces_##S(const OrcParams* p, const R* pos_c, const R* mass_c, R* sr);              \
  /* one whole force evaluation: pmMethodStep (+ short range + correctAccelerations when p3m) */   \
  void orc_force_##S(const OrcParams* p, int p3m, const R* green, const R* pos_c, const R* mass_c, \
                     R* density, R* potential, R* acc);                                            \
  /* run loop pmMethod.cpp:62-135 / p3mMethod.cpp:59-166 with diagnostics (simInfo.cpp:50-127).    \
   * diag = (simLength+1) rows of 12: pe ke px py pz Lx Ly Lz ex ey ez escaped.  Returns the        \
   * number of rows written.  Final state in code units. */                                        \
  int orc_run_##S(const OrcParams* p, int p3m, const float* pos, const float* vel,                 \
                  const float* mass, int simLength, R* diag, R* pos_out, R* vel_out, R* acc_out);

ORC_DECL(float, f32)
ORC_DECL(double, f64)

/* chainingMesh.cpp:60-84: half-shell neighbour list (13 + self), -1 = outside */
void orc_chaining_neighbors(const int* dims, int cell, int* out14);

#ifdef __cplusplus
}
#endif
#endif
