// C-ABI entry points for the dense / implicit-conv GEMM family (SIMT fp32 backend).
// Layouts: activations are row-major [rows, C] with an explicit row stride (NHWC for images),
// Linear weights are [N, K] (torch nn.Linear layout), conv weights are [Cout, R, S, Cin]
// (the physical layout of a channels_last torch Conv2d weight).
#include "gemm_simt.cuh"

using namespace rih;

static inline int is_vec_ok(const void* p, int ld) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0);
}

// y[M,N] = x[M,K] @ w[N,K]^T (+bias) (+residual) (relu) ; reference: torch.nn.Linear (e.g. models/model_attn/gcn.py:92-96)
RIH_API int rih_linear_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy,
                           int M, int N, int K, int relu, int accumulate, const float* res, int ldres,
                           float dropout_p, const unsigned long long* seed_ptr, unsigned long long site, cudaStream_t stream) {
  RIH_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_fwd: bad shape M=%d N=%d K=%d", M, N, K);
  DenseK a{x, ldx, M, is_vec_ok(x, ldx) && (K % 4 == 0)};
  DenseK b{w, ldw, N, is_vec_ok(w, ldw) && (K % 4 == 0)};
  Epilogue ep = make_epilogue(y, ldy, M, N, bias, relu, accumulate ? 1 : 0);
  ep.res = res; ep.ldres = ldres;
  if (dropout_p > 0.f) {
    RIH_REQUIRE(seed_ptr != nullptr && dropout_p < 1.f, "linear_fwd: dropout needs a device seed and p < 1");
    ep.seed_ptr = seed_ptr; ep.site = site; ep.thresh = dropout_thresh(dropout_p); ep.inv_keep = 1.f / (1.f - dropout_p);
  }
  return launch_gemm_simt(a, b, ep, M, N, K, 0, stream, "linear_fwd");
}

// dx[M,K] (+)= dy[M,N] @ w[N,K]
RIH_API int rih_linear_dgrad(const float* dy, int lddy, const float* w, int ldw, float* dx, int lddx,
                             int M, int N, int K, int accumulate, cudaStream_t stream) {
  RIH_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_dgrad: bad shape");
  DenseK a{dy, lddy, M, is_vec_ok(dy, lddy) && (N % 4 == 0)};
  DenseMN b{w, ldw, K, is_vec_ok(w, ldw)};
  Epilogue ep = make_epilogue(dx, lddx, M, K, nullptr, 0, accumulate ? 1 : 0);
  return launch_gemm_simt(a, b, ep, M, K, N, 0, stream, "linear_dgrad");
}

// dw[N,K] (+)= dy[M,N]^T @ x[M,K]
RIH_API int rih_linear_wgrad(const float* dy, int lddy, const float* x, int ldx, float* dw, int lddw,
                             int M, int N, int K, int accumulate, cudaStream_t stream) {
  RIH_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_wgrad: bad shape");
  DenseMN a{dy, lddy, N, is_vec_ok(dy, lddy)};
  DenseMN b{x, ldx, K, is_vec_ok(x, ldx)};
  Epilogue ep = make_epilogue(dw, lddw, N, K, nullptr, 0, accumulate ? 1 : 0);
  return launch_gemm_simt(a, b, ep, N, K, M, 1, stream, "linear_wgrad");
}

static int parse_geom(const int* g, ConvGeom& o) {
  o.N = g[0]; o.H = g[1]; o.W = g[2]; o.Cin = g[3]; o.Ho = g[4]; o.Wo = g[5]; o.Cout = g[6];
  o.R = g[7]; o.S = g[8]; o.stride = g[9]; o.pad = g[10]; o.ldx = g[11]; o.ldy = g[12];
  if (o.N <= 0 || o.H <= 0 || o.W <= 0 || o.Cin <= 0 || o.Cout <= 0 || o.R <= 0 || o.S <= 0 || o.stride <= 0) return 1;
  if (o.Ho != (o.H + 2 * o.pad - o.R) / o.stride + 1 || o.Wo != (o.W + 2 * o.pad - o.S) / o.stride + 1) return 1;
  if (o.ldx < o.Cin || o.ldy < o.Cout) return 1;
  return 0;
}

// NHWC conv forward. geom = {N,H,W,Cin,Ho,Wo,Cout,R,S,stride,pad,ldx,ldy}
// reference: nn.Conv2d call sites models/encoder.py:52,108-116 ; torchvision resnet bottlenecks
RIH_API int rih_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, const int* geom,
                           int relu, cudaStream_t stream) {
  ConvGeom g;
  RIH_REQUIRE(parse_geom(geom, g) == 0, "conv2d_fwd: inconsistent geometry");
  long long M = (long long)g.N * g.Ho * g.Wo;
  int K = g.R * g.S * g.Cin;
  RIH_REQUIRE(M < (1ll << 31), "conv2d_fwd: too many output pixels");
  DenseK b{w, K, g.Cout, is_vec_ok(w, K)};
  Epilogue ep = make_epilogue(y, g.ldy, (int)M, g.Cout, bias, relu, 0);
  if (g.R == 1 && g.S == 1 && g.stride == 1 && g.pad == 0) {
    DenseK a{x, g.ldx, (int)M, is_vec_ok(x, g.ldx) && (K % 4 == 0)};
    return launch_gemm_simt(a, b, ep, (int)M, g.Cout, K, 0, stream, "conv1x1_fwd");
  }
  ConvFwdA a{x, g, (int)M, is_vec_ok(x, g.ldx) && (g.Cin % 4 == 0)};
  return launch_gemm_simt(a, b, ep, (int)M, g.Cout, K, 0, stream, "conv2d_fwd");
}

// dx[N,H,W,Cin] (+)= conv_transpose(dy, w)
RIH_API int rih_conv2d_dgrad(const float* dy, const float* w, float* dx, const int* geom, int accumulate,
                             cudaStream_t stream) {
  ConvGeom g;
  RIH_REQUIRE(parse_geom(geom, g) == 0, "conv2d_dgrad: inconsistent geometry");
  long long M = (long long)g.N * g.H * g.W;
  RIH_REQUIRE(M < (1ll << 31), "conv2d_dgrad: too many pixels");
  int K = g.R * g.S * g.Cout;
  Epilogue ep = make_epilogue(dx, g.ldx, (int)M, g.Cin, nullptr, 0, accumulate ? 1 : 0);
  if (g.R == 1 && g.S == 1 && g.stride == 1 && g.pad == 0) {
    DenseK a{dy, g.ldy, (int)M, is_vec_ok(dy, g.ldy) && (g.Cout % 4 == 0)};
    DenseMN b{w, g.Cin, g.Cin, is_vec_ok(w, g.Cin)};
    return launch_gemm_simt(a, b, ep, (int)M, g.Cin, g.Cout, 0, stream, "conv1x1_dgrad");
  }
  ConvDgradA a{dy, g, (int)M, is_vec_ok(dy, g.ldy) && (g.Cout % 4 == 0)};
  ConvDgradB b{w, g, g.Cin, is_vec_ok(w, g.Cin)};
  return launch_gemm_simt(a, b, ep, (int)M, g.Cin, K, 0, stream, "conv2d_dgrad");
}

// dw[Cout,R,S,Cin] (+)= sum_pixels dy (x) x
RIH_API int rih_conv2d_wgrad(const float* dy, const float* x, float* dw, const int* geom, int accumulate,
                             cudaStream_t stream) {
  ConvGeom g;
  RIH_REQUIRE(parse_geom(geom, g) == 0, "conv2d_wgrad: inconsistent geometry");
  long long P = (long long)g.N * g.Ho * g.Wo;
  RIH_REQUIRE(P < (1ll << 31), "conv2d_wgrad: too many pixels");
  int Kn = g.R * g.S * g.Cin;
  DenseMN a{dy, g.ldy, g.Cout, is_vec_ok(dy, g.ldy)};
  Epilogue ep = make_epilogue(dw, Kn, g.Cout, Kn, nullptr, 0, accumulate ? 1 : 0);
  if (g.R == 1 && g.S == 1 && g.stride == 1 && g.pad == 0) {
    DenseMN b{x, g.ldx, g.Cin, is_vec_ok(x, g.ldx)};
    return launch_gemm_simt(a, b, ep, g.Cout, Kn, (int)P, 1, stream, "conv1x1_wgrad");
  }
  ConvWgradB b{x, g, Kn, is_vec_ok(x, g.ldx) && (g.Cin % 4 == 0)};
  return launch_gemm_simt(a, b, ep, g.Cout, Kn, (int)P, 1, stream, "conv2d_wgrad");
}
