// Host-only check of the slab kernel's item tables and weight-tile packing (csrc/m2h_unet.cu, csrc/unet_slab.cuh):
// replays the kernel's MMA schedule (items -> parts -> shifted slab rows x weight tiles) in plain float on the CPU
// and compares with a direct Conv2d / ConvTranspose2d (k4, s2, p1).  Test infrastructure only; never linked into the product.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../move2hear_active_av_separation_b200/csrc/m2h_unet.cu"

void m2h_set_error(const char*, ...) {}
int m2h_current_device_checked(int*) { return -2; }
int m2h_num_sms(int) { return 0; }

static float frand() { return (float)rand() / RAND_MAX - 0.5f; }

// returns max abs error of the slab schedule vs the direct formula
static double check_deconv(int Cin, int Cout, int H) {
  const int Hp = H + 2, PI = Hp * Hp, B = 2;
  USLayer L; SlabPackDesc D;
  memset(&L, 0, sizeof(L));
  build_deconv_items(Cin, Cout, Cout, Hp, L, D);
  std::vector<float> w((size_t)Cin * Cout * 16), in((size_t)B * PI * Cin, 0.f);
  for (auto& v : w) v = frand();
  for (int b = 0; b < B; ++b) for (int y = 1; y <= H; ++y) for (int x = 1; x <= H; ++x) for (int c = 0; c < Cin; ++c) in[((size_t)(b * PI + y * Hp + x)) * Cin + c] = frand();
  size_t rows = 0;
  for (int i = 0; i < L.n_items; ++i) rows += L.items[i].b_rows;
  if ((int)(rows / Cout) != D.n_blocks) { printf("row/block mismatch %zu %d\n", rows, D.n_blocks); return 1e9; }
  std::vector<float> wt(rows * 64);
  for (size_t i = 0; i < wt.size(); ++i) wt[i] = slab_deconv_w_value(w.data(), D, i);
  const int P = B * PI;
  std::vector<double> acc((size_t)P * 4 * Cout, 0.0);
  std::vector<char> inited((size_t)4 * Cout, 0);
  size_t row0 = 0;
  for (int i = 0; i < L.n_items; ++i) {
    const USItem& it = L.items[i];
    for (int p = 0; p < it.n_parts; ++p) {
      const USPart pt = it.part[p];
      for (int n = 0; n < pt.n; ++n) {
        if (pt.init) { if (inited[pt.tmem_col + n]) { printf("double init\n"); return 1e9; } inited[pt.tmem_col + n] = 1; }
        else if (!inited[pt.tmem_col + n]) { printf("accumulate before init\n"); return 1e9; }
      }
      for (int q = 0; q < P; ++q) {
        const int src = q - (Hp + 1) + it.a_shift;      // slab row m + a_shift <-> position q0 - (Wp+1) + m + a_shift
        if (src < 0 || src >= P) continue;
        for (int n = 0; n < pt.n; ++n) {
          double s = 0;
          for (int k = 0; k < it.n_k16 * 16; ++k)
            s += (double)in[(size_t)src * Cin + it.chunk * 64 + it.a_k16 * 16 + k] * wt[(row0 + pt.b_row + n) * 64 + k];
          acc[(size_t)q * 4 * Cout + pt.tmem_col + n] += s;
        }
      }
    }
    row0 += it.b_rows;
  }
  double err = 0;
  for (int b = 0; b < B; ++b) for (int oy = 0; oy < 2 * H; ++oy) for (int ox = 0; ox < 2 * H; ++ox) for (int co = 0; co < Cout; ++co) {
    double ref = 0;      // ConvTranspose2d: out[oy] += in[iy] * w[kh], oy = 2*iy - 1 + kh
    for (int kh = 0; kh < 4; ++kh) for (int kw = 0; kw < 4; ++kw) {
      const int ty = oy + 1 - kh, tx = ox + 1 - kw;
      if (ty < 0 || tx < 0 || (ty & 1) || (tx & 1)) continue;
      const int iy = ty / 2, ix = tx / 2;
      if (iy >= H || ix >= H) continue;
      for (int c = 0; c < Cin; ++c) ref += (double)in[((size_t)(b * PI + (iy + 1) * Hp + ix + 1)) * Cin + c] * w[(((size_t)c * Cout + co) * 4 + kh) * 4 + kw];
    }
    const int y = oy / 2 + 1, x = ox / 2 + 1, ph = (oy & 1) * 2 + (ox & 1);
    const double got = acc[((size_t)(b * PI + y * Hp + x)) * 4 * Cout + ph * Cout + co];
    if (fabs(got - ref) > err) err = fabs(got - ref);
  }
  return err;
}

static double check_conv0() {
  const int Hc = 16, Hp = 18, PI = Hp * Hp, B = 2, Cin = 32, CinFull = 33, Cout = 64;
  USLayer L; SlabPackDesc D;
  memset(&L, 0, sizeof(L));
  build_conv0_items(Hp, CinFull, L, D);
  std::vector<float> w((size_t)Cout * CinFull * 16), img((size_t)B * 32 * 32 * Cin), in((size_t)B * PI * 128, 0.f);
  for (auto& v : w) v = frand();
  for (auto& v : img) v = frand();
  for (int b = 0; b < B; ++b) for (int h = 0; h < 32; ++h) for (int x = 0; x < 32; ++x) for (int c = 0; c < Cin; ++c)
    in[((size_t)(b * PI + (h / 2 + 1) * Hp + x / 2 + 1)) * 128 + ((h & 1) * 2 + (x & 1)) * 32 + c] = img[(((size_t)b * 32 + h) * 32 + x) * Cin + c];
  size_t rows = 0;
  for (int i = 0; i < L.n_items; ++i) rows += L.items[i].b_rows;
  std::vector<float> wt(rows * 64);
  for (size_t i = 0; i < wt.size(); ++i) wt[i] = slab_conv_w_value(w.data(), D, i);
  const int P = B * PI;
  std::vector<double> acc((size_t)P * Cout, 0.0);
  size_t row0 = 0;
  for (int i = 0; i < L.n_items; ++i) {
    const USItem& it = L.items[i];
    const USPart pt = it.part[0];
    for (int q = 0; q < P; ++q) {
      const int src = q - (Hp + 1) + it.a_shift;
      if (src < 0 || src >= P) continue;
      for (int n = 0; n < pt.n; ++n) {
        double s = 0;
        for (int k = 0; k < it.n_k16 * 16; ++k) s += (double)in[(size_t)src * 128 + it.chunk * 64 + it.a_k16 * 16 + k] * wt[(row0 + pt.b_row + n) * 64 + k];
        acc[(size_t)q * Cout + n] += s;
      }
    }
    row0 += it.b_rows;
  }
  double err = 0;
  for (int b = 0; b < B; ++b) for (int oy = 0; oy < Hc; ++oy) for (int ox = 0; ox < Hc; ++ox) for (int co = 0; co < Cout; ++co) {
    double ref = 0;
    for (int kh = 0; kh < 4; ++kh) for (int kw = 0; kw < 4; ++kw) {
      const int iy = 2 * oy + kh - 1, ix = 2 * ox + kw - 1;
      if (iy < 0 || iy >= 32 || ix < 0 || ix >= 32) continue;
      for (int c = 0; c < Cin; ++c) ref += (double)img[(((size_t)b * 32 + iy) * 32 + ix) * Cin + c] * w[(((size_t)co * CinFull + c) * 4 + kh) * 4 + kw];
    }
    const double got = acc[((size_t)(b * PI + (oy + 1) * Hp + ox + 1)) * Cout + co];
    if (fabs(got - ref) > err) err = fabs(got - ref);
  }
  return err;
}

int main() {
  srand(1);
  const double e4 = check_deconv(128, 32, 16), e3 = check_deconv(256, 64, 8), e2 = check_deconv(128, 64, 4), e0 = check_conv0();
  printf("deconv(128->32,16x16) %.3e  deconv(256->64,8x8) %.3e  deconv(128->64,4x4) %.3e  conv0 %.3e\n", e4, e3, e2, e0);
  return (e4 < 1e-4 && e3 < 1e-4 && e2 < 1e-4 && e0 < 1e-4) ? 0 : 1;
}
