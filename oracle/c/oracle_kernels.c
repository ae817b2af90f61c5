/* CPU restatement in plain C of the integer/index steps of the hot path and of the direct convolution,
 * with double-precision accumulation -- TEST INFRASTRUCTURE (checker only; never linked by the product).
 *
 *   vq_nearest_ref     VectorQuantizer.forward distance + argmin   /root/reference/basicsr/archs/vqgan_arch.py:38-44
 *   argmax_lookup_ref  softmax->topk(1)->one-hot@E                 codeformer_arch.py:257-259, vqgan_arch.py:72-84
 *   conv2d_nhwc_ref    nn.Conv2d 3x3/1x1 (+Downsample/Upsample)    vqgan_arch.py:117-138,147-151
 *   group_norm_ref     GroupNorm(32,C,eps)                         vqgan_arch.py:14-15
 * "parity unpinned" by the reference (it ships no vectors); pinned by tests/golden (oracle/gen_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* z [T,D], E [K,D] -> idx [T] = first index of the minimum of |z|^2+|e|^2-2 z.e (evaluated in double) */
void vq_nearest_ref(const float* z, const float* E, int T, int D, int K, int64_t* idx, double* min_gap) {
  double gap = INFINITY;
  for (int t = 0; t < T; ++t) {
    double best = INFINITY, second = INFINITY;
    int bi = 0;
    double z2 = 0;
    for (int d = 0; d < D; ++d) z2 += (double)z[(size_t)t * D + d] * z[(size_t)t * D + d];
    for (int k = 0; k < K; ++k) {
      double e2 = 0, ze = 0;
      for (int d = 0; d < D; ++d) {
        const double e = E[(size_t)k * D + d];
        e2 += e * e;
        ze += e * (double)z[(size_t)t * D + d];
      }
      const double dist = z2 + e2 - 2 * ze;
      if (dist < best) { second = best; best = dist; bi = k; }
      else if (dist < second) second = dist;
    }
    idx[t] = bi;
    if (second - best < gap) gap = second - best;
  }
  if (min_gap) *min_gap = gap;
}

/* logits [T,K] -> idx (first maximum), quant [T,D] = E[idx] */
void argmax_lookup_ref(const float* logits, const float* E, int T, int K, int D, int64_t* idx, float* quant) {
  for (int t = 0; t < T; ++t) {
    int bi = 0;
    float best = logits[(size_t)t * K];
    for (int k = 1; k < K; ++k)
      if (logits[(size_t)t * K + k] > best) { best = logits[(size_t)t * K + k]; bi = k; }
    idx[t] = bi;
    for (int d = 0; d < D; ++d) quant[(size_t)t * D + d] = E[(size_t)bi * D + d];
  }
}

/* in [N,H,W,Cin] NHWC, w OIHW [Cout,Cin,k,k], out [N,Ho,Wo,Cout]; mode 0 same, 1 down (pad r/b, s2), 2 up (nearest x2) */
void conv2d_nhwc_ref(const float* in, const float* w, const float* bias, float* out, int N, int H, int W, int Cin,
                     int Cout, int k, int mode) {
  const int Ho = mode == 1 ? H / 2 : (mode == 2 ? H * 2 : H), Wo = mode == 1 ? W / 2 : (mode == 2 ? W * 2 : W);
  for (int n = 0; n < N; ++n)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox)
        for (int co = 0; co < Cout; ++co) {
          double acc = bias ? bias[co] : 0.0;
          for (int r = 0; r < k; ++r)
            for (int s = 0; s < k; ++s) {
              int iy, ix;
              if (mode == 1) { iy = oy * 2 + r; ix = ox * 2 + s; if (iy >= H || ix >= W) continue; }
              else if (mode == 2) {
                iy = oy + r - 1; ix = ox + s - 1;
                if (iy < 0 || ix < 0 || iy >= 2 * H || ix >= 2 * W) continue;
                iy >>= 1; ix >>= 1;
              } else {
                iy = oy + r - k / 2; ix = ox + s - k / 2;
                if (iy < 0 || ix < 0 || iy >= H || ix >= W) continue;
              }
              const float* ip = in + (((size_t)n * H + iy) * W + ix) * Cin;
              for (int ci = 0; ci < Cin; ++ci) acc += (double)ip[ci] * w[(((size_t)co * Cin + ci) * k + r) * k + s];
            }
          out[(((size_t)n * Ho + oy) * Wo + ox) * Cout + co] = (float)acc;
        }
}

/* x [N,HW,C] NHWC -> y = (x-mean_g)*rstd_g*gamma+beta, biased variance per (n, group) */
void group_norm_ref(const float* x, const float* gamma, const float* beta, float* y, int N, int HW, int C, int groups,
                    double eps) {
  const int cpg = C / groups;
  for (int n = 0; n < N; ++n)
    for (int g = 0; g < groups; ++g) {
      double s = 0, q = 0;
      for (int p = 0; p < HW; ++p)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { const double v = x[((size_t)n * HW + p) * C + c]; s += v; q += v * v; }
      const double cnt = (double)HW * cpg, mean = s / cnt, var = q / cnt - mean * mean, rstd = 1.0 / sqrt(var + eps);
      for (int p = 0; p < HW; ++p)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          const size_t i = ((size_t)n * HW + p) * C + c;
          y[i] = (float)(((double)x[i] - mean) * rstd * gamma[c] + beta[c]);
        }
    }
}
