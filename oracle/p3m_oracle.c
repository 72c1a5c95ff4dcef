/* oracle/p3m_oracle.c -- TEST INFRASTRUCTURE ONLY (see p3m_oracle.h).
 * Instantiates the restatement in p3m_oracle_impl.h for float (the reference's precision) and for
 * double (same algorithm, used for the 1e-6 parity bar). */
#include "p3m_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* chainingMesh.cpp:60-84 getNeighborsAndSelf + tripleToFlatIndex */
static int tri(const int* M, int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0 || x >= M[0] || y >= M[1] || z >= M[2]) return -1;
  return x + y * M[0] + z * M[0] * M[1];
}

void orc_chaining_neighbors(const int* M, int cell, int* nb) {
  int cx = cell % M[0], cy = (cell / M[0]) % M[1], cz = cell / (M[0] * M[1]);
  int i = 0;
  for (int t = -1; t <= 1; ++t)
    for (int s = -1; s <= 1; ++s) nb[i++] = tri(M, cx + t, cy - 1, cz + s);
  for (int s = -1; s <= 1; ++s) nb[i++] = tri(M, cx + s, cy, cz - 1);
  nb[12] = tri(M, cx - 1, cy, cz);
  nb[13] = cell;
}

/* std::numbers::pi_v<float> / pi_v<double> */
#define R float
#define S f32
#define PI_R 3.14159265358979323846f
#define SIN sinf
#define COS cosf
#define POW powf
#define SQRT sqrtf
#define ROUND roundf
#include "p3m_oracle_impl.h"
#undef R
#undef S
#undef PI_R
#undef SIN
#undef COS
#undef POW
#undef SQRT
#undef ROUND

#define R double
#define S f64
#define PI_R 3.14159265358979323846
#define SIN sin
#define COS cos
#define POW pow
#define SQRT sqrt
#define ROUND round
#include "p3m_oracle_impl.h"
