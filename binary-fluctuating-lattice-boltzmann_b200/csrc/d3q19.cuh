// D3Q19 lattice constants and the Duenweg-Schiller-Ladd moment transforms, device side.
// Semantics follow the reference (velocity order LBM_d3q19.H:12-32, weights :34-54, mode norms :56-76,
// forward transform :100-156, inverse :167-247); the arithmetic is re-derived for the GPU
// (multiplications by reciprocals instead of divisions, FMA-friendly grouping).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bflbm {

constexpr int Q = 19;
constexpr double CS2 = 1.0 / 3.0;

// c_i, compile-time so that every use is folded into immediates / address offsets
__host__ __device__ constexpr int cx(int i) {
  constexpr int v[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
  return v[i];
}
__host__ __device__ constexpr int cy(int i) {
  constexpr int v[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 1, -1, 1, -1, 0, 0, 0, 0};
  return v[i];
}
__host__ __device__ constexpr int cz(int i) {
  constexpr int v[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1};
  return v[i];
}
__host__ __device__ constexpr double wq(int i) { return i == 0 ? 1.0 / 3.0 : (i < 7 ? 1.0 / 18.0 : 1.0 / 36.0); }
// b_k = sum_i w_i e_ki^2
__host__ __device__ constexpr double bnorm(int k) {
  constexpr double v[Q] = {1.0,       1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0, 4.0 / 3.0, 4.0 / 9.0,
                           1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 2.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0, 2.0 / 9.0,
                           2.0 / 9.0, 2.0 / 9.0, 2.0,       4.0 / 3.0, 4.0 / 9.0};
  return v[k];
}
// sqrt(b_k) for the noise amplitudes of modes 4..18
__host__ __device__ constexpr double sqrt_bnorm(int k) {
  constexpr double v[Q] = {1.0,
                           0.57735026918962576451, 0.57735026918962576451, 0.57735026918962576451,
                           0.81649658092772603273, 1.15470053837925152902, 0.66666666666666666667,
                           0.33333333333333333333, 0.33333333333333333333, 0.33333333333333333333,
                           0.81649658092772603273, 0.81649658092772603273, 0.81649658092772603273,
                           0.47140452079103168293, 0.47140452079103168293, 0.47140452079103168293,
                           1.41421356237309504880, 1.15470053837925152902, 0.66666666666666666667};
  return v[k];
}

// m = M f
__device__ __forceinline__ void moments(const double (&f)[Q], double (&m)[Q]) {
  const double mx1 = f[1] - f[2], mxx1 = f[1] + f[2];
  const double my1 = f[3] - f[4], myy1 = f[3] + f[4];
  const double mz1 = f[5] - f[6], mzz1 = f[5] + f[6];
  // xy diagonals 7..10
  const double a78 = f[7] + f[8], s78 = f[7] - f[8], a910 = f[9] + f[10], s910 = f[9] - f[10];
  const double mx2 = s78 + s910, my3 = s78 - s910, mxy = a78 - a910, mxx2 = a78 + a910;
  // yz diagonals 11..14
  const double a1112 = f[11] + f[12], s1112 = f[11] - f[12], a1314 = f[13] + f[14], s1314 = f[13] - f[14];
  const double my2 = s1112 + s1314, mz3 = s1112 - s1314, myz = a1112 - a1314, myy2 = a1112 + a1314;
  // xz diagonals 15..18 : 15 (+,+) 16 (-,-) 17 (+x,-z) 18 (-x,+z)
  const double a1516 = f[15] + f[16], s1516 = f[15] - f[16], a1718 = f[17] + f[18], s1718 = f[17] - f[18];
  const double mz2 = s1516 - s1718, mx3 = s1516 + s1718, mxz = a1516 - a1718, mzz2 = a1516 + a1718;

  const double mc0 = f[0];
  const double mc1 = mxx1 + myy1 + mzz1;
  const double mc2 = mxx2 + myy2 + mzz2;
  m[0] = mc0 + mc1 + mc2;
  m[1] = mx1 + mx2 + mx3;
  m[2] = my1 + my2 + my3;
  m[3] = mz1 + mz2 + mz3;
  m[4] = mc2 - mc0;
  m[5] = 3. * mxx1 - mc1 + mc2 - 3. * myy2;
  m[6] = myy1 - mzz1 + mxx2 - mzz2;
  m[7] = mxy;
  m[8] = myz;
  m[9] = mxz;
  m[10] = m[1] - 3. * mx1;
  m[11] = m[2] - 3. * my1;
  m[12] = m[3] - 3. * mz1;
  m[13] = mx2 - mx3;
  m[14] = my2 - my3;
  m[15] = mz2 - mz3;
  m[16] = m[0] - 3. * mc1;
  m[17] = mc1 - 3. * mxx1 + mc2 - 3. * myy2;
  m[18] = mzz1 - myy1 + mxx2 - mzz2;
}

// f = M^-1 m,  M^-1[i,k] = w_i e_ki / b_k
__device__ __forceinline__ void populations(const double (&mom)[Q], double (&f)[Q]) {
  const double m0 = mom[0] * (1. / 36.), m1 = mom[1] * (1. / 12.), m2 = mom[2] * (1. / 12.), m3 = mom[3] * (1. / 12.);
  const double m4 = mom[4] * (1. / 24.), m5 = mom[5] * (1. / 48.), m6 = mom[6] * (1. / 16.);
  const double m7 = mom[7] * 0.25, m8 = mom[8] * 0.25, m9 = mom[9] * 0.25;
  const double m10 = mom[10] * (1. / 24.), m11 = mom[11] * (1. / 24.), m12 = mom[12] * (1. / 24.);
  const double m13 = mom[13] * 0.125, m14 = mom[14] * 0.125, m15 = mom[15] * 0.125;
  const double m16 = mom[16] * (1. / 72.), m17 = mom[17] * (1. / 48.), m18 = mom[18] * (1. / 16.);

  const double mc0 = 12. * (m0 - m4 + m16);
  const double mc1 = 2. * (m0 - 2. * m16);
  const double mc2 = m0 + m4 + m16;
  const double mx1 = 2. * (m1 - 2. * m10), my1 = 2. * (m2 - 2. * m11), mz1 = 2. * (m3 - 2. * m12);
  const double mx2 = m1 + m10 + m13, my2 = m2 + m11 + m14, mz2 = m3 + m12 + m15;
  const double mx3 = m1 + m10 - m13, my3 = m2 + m11 - m14, mz3 = m3 + m12 - m15;
  const double mxx1 = mc1 + 4. * (m5 - m17);
  const double myy1 = mc1 - 2. * (m5 - m6) + 2. * (m17 - m18);
  const double mzz1 = mc1 - 2. * (m5 + m6) + 2. * (m17 + m18);
  const double mxy2 = mc2 + (m5 + m6) + (m17 + m18);
  const double mxz2 = mc2 + (m5 - m6) + (m17 - m18);
  const double myz2 = mc2 - 2. * (m5 + m17);

  f[0] = mc0;
  f[1] = mxx1 + mx1;
  f[2] = mxx1 - mx1;
  f[3] = myy1 + my1;
  f[4] = myy1 - my1;
  f[5] = mzz1 + mz1;
  f[6] = mzz1 - mz1;
  const double p7 = mxy2 + m7, q7 = mxy2 - m7, r7 = mx2 + my3, t7 = mx2 - my3;
  f[7] = p7 + r7;
  f[8] = p7 - r7;
  f[9] = q7 + t7;
  f[10] = q7 - t7;
  const double p11 = myz2 + m8, q11 = myz2 - m8, r11 = my2 + mz3, t11 = my2 - mz3;
  f[11] = p11 + r11;
  f[12] = p11 - r11;
  f[13] = q11 + t11;
  f[14] = q11 - t11;
  const double p15 = mxz2 + m9, q15 = mxz2 - m9, r15 = mz2 + mx3, t15 = mx3 - mz2;
  f[15] = p15 + r15;
  f[16] = p15 - r15;
  f[17] = q15 + t15;
  f[18] = q15 - t15;
}

}  // namespace bflbm
