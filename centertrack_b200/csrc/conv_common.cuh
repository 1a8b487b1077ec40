// Device-side pieces shared by the SIMT (fp32-exact) and tcgen05 (bf16) implicit-GEMM engines:
// output-pixel decomposition, conv window addressing and DCNv2 bilinear sampling.
#pragma once
#include "common.cuh"

namespace ctb {

struct ConvGeom {
  int B, H, W, C_in, ld_in;
  int C_out, KH, KW, stride, pad, pad_w, OH, OW;
  int ld_out, out_mode, relu, ld_res, head_act, sig_from;
  float depth_scale;
  int ld_om;
  int P_out;      // B*OH*OW
  int K_total;    // KH*KW*C_in
};

inline ConvGeom make_geom(const ct_conv_desc* d) {
  ConvGeom g;
  g.B = d->B; g.H = d->H; g.W = d->W; g.C_in = d->C_in; g.ld_in = d->ld_in;
  g.C_out = d->C_out; g.KH = d->KH; g.KW = d->KW; g.stride = d->stride; g.pad = d->pad;
  g.pad_w = d->pad_w1 > 0 ? d->pad_w1 - 1 : d->pad;
  g.OH = d->OH; g.OW = d->OW; g.ld_out = d->ld_out; g.out_mode = d->out_mode;
  g.relu = d->relu; g.ld_res = d->ld_res; g.head_act = d->head_act; g.sig_from = d->sig_from;
  g.depth_scale = d->depth_scale; g.ld_om = d->ld_om;
  g.P_out = d->B * d->OH * d->OW;
  g.K_total = d->KH * d->KW * d->C_in;
  return g;
}

// Modulated-deformable sampling parameters of one (output pixel, tap):  SURVEY.md Appendix B.
//   py = y - 1 + i + dy_k,  px = x - 1 + j + dx_k ;  S = 0 unless -1 < py < H and -1 < px < W ;
//   bilinear over the integer neighbours that lie inside the image; the whole sample times mask.
struct DcnTap {
  int off[4];     // element offsets (relative to the image base) of the 4 corner pixels
  float w[4];     // bilinear weight x mask, 0 for corners outside the image
};

__device__ __forceinline__ DcnTap dcn_tap(const float* __restrict__ om_px, int tap, int oy, int ox,
                                          int H, int W, int ld_in) {
  DcnTap t;
  const float dy = om_px[2 * tap], dx = om_px[2 * tap + 1], m = om_px[18 + tap];
  const float py = (float)(oy - 1 + tap / 3) + dy;
  const float px = (float)(ox - 1 + tap % 3) + dx;
  const bool valid = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
  const float y0f = floorf(py), x0f = floorf(px);
  const float ly = py - y0f, lx = px - x0f, hy = 1.f - ly, hx = 1.f - lx;
  const int y0 = (int)y0f, x0 = (int)x0f;
  const bool y0ok = valid && y0 >= 0 && y0 <= H - 1, y1ok = valid && y0 + 1 >= 0 && y0 + 1 <= H - 1;
  const bool x0ok = x0 >= 0 && x0 <= W - 1, x1ok = x0 + 1 >= 0 && x0 + 1 <= W - 1;
  const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
  const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
  t.off[0] = (yc0 * W + xc0) * ld_in; t.w[0] = (y0ok && x0ok) ? hy * hx * m : 0.f;
  t.off[1] = (yc0 * W + xc1) * ld_in; t.w[1] = (y0ok && x1ok) ? hy * lx * m : 0.f;
  t.off[2] = (yc1 * W + xc0) * ld_in; t.w[2] = (y1ok && x0ok) ? ly * hx * m : 0.f;
  t.off[3] = (yc1 * W + xc1) * ld_in; t.w[3] = (y1ok && x1ok) ? ly * lx * m : 0.f;
  return t;
}

// Apply the per-channel epilogue transform of fp32 head planes.
__device__ __forceinline__ float head_transform(float v, int head_act, float depth_scale) {
  if (head_act == CT_HEAD_SIGMOID) return sigmoidf_ref(v);
  if (head_act == CT_HEAD_DEPTH) return (1.f / (sigmoidf_ref(v) + 1e-6f) - 1.f) * depth_scale;
  return v;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef int (*TmapEncodeFnRaw)(void*, int, unsigned, void*, const unsigned long long*, const unsigned long long*,
                               const unsigned*, const unsigned*, int, int, int, int);
void* tmap_encode_raw();

int conv_forward_simt(const ct_conv_desc* d, cudaStream_t st);
int conv_forward_tc(const ct_conv_desc* d, cudaStream_t st);
int conv_forward_halo(const ct_conv_desc* d, cudaStream_t st);
int halo_blocks(int C_in, int KH, int KW);
int halo_set_trace(void* buf);
int tc_set_trace(void* buf);
int halo_set_watch(void* mapped_host_buf);

}  // namespace ctb
