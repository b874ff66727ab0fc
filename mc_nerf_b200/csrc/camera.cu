// Camera model: learnable intrinsics and se(3) extrinsics, forward and analytic backward.
// One thread per camera (110 cameras: latency-bound, a single small launch replaces the reference's
// ~150 tiny ATen launches + a 110-iteration torch.inverse loop).
// ref: model/mc_nerf.py:171-186 (add_weights2intr), :204-210 (inverse_intrinsic), :269-316 (se3_to_SE3).
#include "common.cuh"

namespace {

__device__ __forceinline__ void intrinsics_fwd_dev(int i, const float* wfx, const float* wfy,
                                 const float* wux, const float* wuy, int n, float H, float W,
                                 float* K, float* Kinv) {
  // sic: fy is scaled by the image WIDTH in the reference (model/mc_nerf.py:173)
  float fx = fabsf(W * wfx[i]), fy = fabsf(W * wfy[i]);
  float ux = fabsf(W * 0.5f * wux[i]), uy = fabsf(H * 0.5f * wuy[i]);
  float* k = K + 9 * i;
  k[0] = fx; k[1] = 0.f; k[2] = ux;
  k[3] = 0.f; k[4] = fy; k[5] = uy;
  k[6] = 0.f; k[7] = 0.f; k[8] = 1.f;
  if (Kinv) {
    float ifx = 1.f / fx, ify = 1.f / fy;
    float* q = Kinv + 9 * i;
    q[0] = ifx; q[1] = 0.f; q[2] = -ux * ifx;
    q[3] = 0.f; q[4] = ify; q[5] = -uy * ify;
    q[6] = 0.f; q[7] = 0.f; q[8] = 1.f;
  }
}

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ void intrinsics_bwd_dev(int i, const float* wfx, const float* wfy,
                                 const float* wux, const float* wuy, int n, float H, float W,
                                 const float* gK, const float* gKinv,
                                 float* gfx, float* gfy, float* gux,
                                 float* guy) {
  float afx = W * wfx[i], afy = W * wfy[i], aux = W * 0.5f * wux[i], auy = H * 0.5f * wuy[i];
  float fx = fabsf(afx), fy = fabsf(afy), ux = fabsf(aux), uy = fabsf(auy);
  float d_fx = 0.f, d_fy = 0.f, d_ux = 0.f, d_uy = 0.f;
  if (gK) {
    const float* g = gK + 9 * i;
    d_fx += g[0]; d_ux += g[2]; d_fy += g[4]; d_uy += g[5];
  }
  if (gKinv) {
    const float* g = gKinv + 9 * i;
    float ifx = 1.f / fx, ify = 1.f / fy;
    // Kinv00 = 1/fx, Kinv02 = -ux/fx, Kinv11 = 1/fy, Kinv12 = -uy/fy
    d_fx += -g[0] * ifx * ifx + g[2] * ux * ifx * ifx;
    d_ux += -g[2] * ifx;
    d_fy += -g[4] * ify * ify + g[5] * uy * ify * ify;
    d_uy += -g[5] * ify;
  }
  gfx[i] = d_fx * sgn(afx) * W;
  gfy[i] = d_fy * sgn(afy) * W;
  gux[i] = d_ux * sgn(aux) * W * 0.5f;
  guy[i] = d_uy * sgn(auy) * H * 0.5f;
}

// 11-term series (nth = 10) of A = sin t / t, B = (1 - cos t)/t^2, C = (t - sin t)/t^3 and their
// derivatives with respect to t.  The reference evaluates exactly these truncated series, so closed
// forms would DIVERGE from it for large angles.
struct Series { float A, B, C, dA, dB, dC; };
__device__ Series taylor_abc(float t) {
  Series s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float t2 = t * t;
  float p = 1.f;       // t^(2i)
  float pm1 = 0.f;     // t^(2i-1)
  double dA = 1.0, dB = 1.0, dC = 1.0;
  for (int i = 0; i <= 10; ++i) {
    if (i > 0) dA *= (double)((2 * i) * (2 * i + 1));
    dB *= (double)((2 * i + 1) * (2 * i + 2));
    dC *= (double)((2 * i + 2) * (2 * i + 3));
    float sg = (i & 1) ? -1.f : 1.f;
    s.A += sg * p / (float)dA;
    s.B += sg * p / (float)dB;
    s.C += sg * p / (float)dC;
    if (i > 0) {
      float c = sg * (float)(2 * i) * pm1;
      s.dA += c / (float)dA;
      s.dB += c / (float)dB;
      s.dC += c / (float)dC;
    }
    pm1 = p * t;       // t^(2i+1)
    p = p * t2;
  }
  return s;
}

__device__ __forceinline__ void skew(const float w[3], float wx[9]) {
  wx[0] = 0.f;   wx[1] = -w[2]; wx[2] = w[1];
  wx[3] = w[2];  wx[4] = 0.f;   wx[5] = -w[0];
  wx[6] = -w[1]; wx[7] = w[0];  wx[8] = 0.f;
}
__device__ __forceinline__ void mat3_mul(const float a[9], const float b[9], float c[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

__device__ __forceinline__ void se3_fwd_dev(int i, const float* wu, int n, float* Rt) {
  float w[3] = {wu[6 * i], wu[6 * i + 1], wu[6 * i + 2]};
  float u[3] = {wu[6 * i + 3], wu[6 * i + 4], wu[6 * i + 5]};
  float th = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  Series s = taylor_abc(th);
  float wx[9], wx2[9];
  skew(w, wx);
  mat3_mul(wx, wx, wx2);
  float* o = Rt + 12 * i;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float t = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float I = (r == c) ? 1.f : 0.f;
      o[4 * r + c] = I + s.A * wx[3 * r + c] + s.B * wx2[3 * r + c];
      float V = I + s.B * wx[3 * r + c] + s.C * wx2[3 * r + c];
      t += V * u[c];
    }
    o[4 * r + 3] = t;
  }
}

__device__ __forceinline__ void se3_bwd_dev(int i, const float* wu, const float* gRt, int n, float* gwu) {
  float w[3] = {wu[6 * i], wu[6 * i + 1], wu[6 * i + 2]};
  float u[3] = {wu[6 * i + 3], wu[6 * i + 4], wu[6 * i + 5]};
  float th = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  Series s = taylor_abc(th);
  float wx[9], wx2[9];
  skew(w, wx);
  mat3_mul(wx, wx, wx2);
  const float* g = gRt + 12 * i;
  float gR[9], gt[3], gV[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gR[3 * r + c] = g[4 * r + c];
    gt[r] = g[4 * r + 3];
  }
  // t = V u  ->  gu = V^T gt, gV = gt u^T
  float gu[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float V = ((r == c) ? 1.f : 0.f) + s.B * wx[3 * r + c] + s.C * wx2[3 * r + c];
      gu[c] += V * gt[r];
      gV[3 * r + c] = gt[r] * u[c];
    }
  float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    gA += gR[k] * wx[k];
    gB += gR[k] * wx2[k] + gV[k] * wx[k];
    gC += gV[k] * wx2[k];
  }
  // d<G, wx^2>/d wx = G wx^T + wx^T G
  float P[9];   // P = B*gR + C*gV  (the coefficient matrix of wx^2)
  float G[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    P[k] = s.B * gR[k] + s.C * gV[k];
    G[k] = s.A * gR[k] + s.B * gV[k];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) a += P[3 * r + k] * wx[3 * c + k] + wx[3 * k + r] * P[3 * k + c];
      G[3 * r + c] += a;
    }
  float gw[3] = {G[7] - G[5], G[2] - G[6], G[3] - G[1]};
  float gth = gA * s.dA + gB * s.dB + gC * s.dC;
  // theta = |w| (norm backward; NaN at w = 0 exactly as torch's norm backward gives 0/0 -> we return 0 there)
  float inv = th > 0.f ? 1.f / th : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gwu[6 * i + k] = gw[k] + gth * w[k] * inv;
    gwu[6 * i + 3 + k] = gu[k];
  }
}

// Calibration-point reprojection px = K [R|t] X / z, one thread per camera (P points each).
// ref: model/mc_nerf.py:147-152, 236-241, 260-267.
__device__ __forceinline__ void reproject_fwd_dev(int c, const float* wpts, const float* K,
                                const float* Rt, int n, int P, float* pix) {
  const float* k = K + 9 * c;
  const float* m = Rt + 12 * c;
  for (int p = 0; p < P; ++p) {
    const float* X = wpts + ((size_t)c * P + p) * 3;
    float xc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) xc[r] = m[4 * r] * X[0] + m[4 * r + 1] * X[1] + m[4 * r + 2] * X[2] + m[4 * r + 3];
    float h[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) h[r] = k[3 * r] * xc[0] + k[3 * r + 1] * xc[1] + k[3 * r + 2] * xc[2];
    pix[((size_t)c * P + p) * 2] = h[0] / h[2];
    pix[((size_t)c * P + p) * 2 + 1] = h[1] / h[2];
  }
}

__device__ __forceinline__ void reproject_bwd_dev(int c, const float* wpts, const float* K,
                                const float* Rt, const float* gpix, int n, int P,
                                float* gK, float* gRt) {
  const float* k = K + 9 * c;
  const float* m = Rt + 12 * c;
  float dk[9], dm[12];
#pragma unroll
  for (int i = 0; i < 9; ++i) dk[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) dm[i] = 0.f;
  for (int p = 0; p < P; ++p) {
    const float* X = wpts + ((size_t)c * P + p) * 3;
    float xc[3], h[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) xc[r] = m[4 * r] * X[0] + m[4 * r + 1] * X[1] + m[4 * r + 2] * X[2] + m[4 * r + 3];
#pragma unroll
    for (int r = 0; r < 3; ++r) h[r] = k[3 * r] * xc[0] + k[3 * r + 1] * xc[1] + k[3 * r + 2] * xc[2];
    const float gx = gpix[((size_t)c * P + p) * 2], gy = gpix[((size_t)c * P + p) * 2 + 1];
    const float iz = 1.f / h[2];
    const float gh[3] = {gx * iz, gy * iz, -(gx * h[0] + gy * h[1]) * iz * iz};
    float gxc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        dk[3 * r + j] += gh[r] * xc[j];
        gxc[j] += k[3 * r + j] * gh[r];
      }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      dm[4 * r] += gxc[r] * X[0];
      dm[4 * r + 1] += gxc[r] * X[1];
      dm[4 * r + 2] += gxc[r] * X[2];
      dm[4 * r + 3] += gxc[r];
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) gK[9 * c + i] = dk[i];
#pragma unroll
  for (int i = 0; i < 12; ++i) gRt[12 * c + i] = dm[i];
}

__global__ void intrinsics_fwd_k(const float* __restrict__ wfx, const float* __restrict__ wfy, const float* __restrict__ wux, const float* __restrict__ wuy, int n, float H, float W, float* __restrict__ K, float* __restrict__ Kinv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) intrinsics_fwd_dev(i, wfx, wfy, wux, wuy, n, H, W, K, Kinv);
}
__global__ void intrinsics_bwd_k(const float* __restrict__ wfx, const float* __restrict__ wfy, const float* __restrict__ wux, const float* __restrict__ wuy, int n, float H, float W, const float* __restrict__ gK, const float* __restrict__ gKinv, float* __restrict__ gfx, float* __restrict__ gfy, float* __restrict__ gux, float* __restrict__ guy) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) intrinsics_bwd_dev(i, wfx, wfy, wux, wuy, n, H, W, gK, gKinv, gfx, gfy, gux, guy);
}
__global__ void se3_fwd_k(const float* __restrict__ wu, int n, float* __restrict__ Rt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) se3_fwd_dev(i, wu, n, Rt);
}
__global__ void se3_bwd_k(const float* __restrict__ wu, const float* __restrict__ gRt, int n, float* __restrict__ gwu) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) se3_bwd_dev(i, wu, gRt, n, gwu);
}
__global__ void reproject_fwd_k(const float* __restrict__ wpts, const float* __restrict__ K, const float* __restrict__ Rt, int n, int P, float* __restrict__ pix) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reproject_fwd_dev(i, wpts, K, Rt, n, P, pix);
}
__global__ void reproject_bwd_k(const float* __restrict__ wpts, const float* __restrict__ K, const float* __restrict__ Rt, const float* __restrict__ gpix, int n, int P, float* __restrict__ gK, float* __restrict__ gRt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reproject_bwd_dev(i, wpts, K, Rt, gpix, n, P, gK, gRt);
}

// ---- the whole camera model of a train step in ONE launch each way (one thread per camera):
// forward : intrinsics (K, K^-1) + se3 -> SE3 of the main and of the calibration poses + reprojection of the calibration
//           points through the calibration pose (ref: model/mc_nerf.py:75-76 = add_weights2param + get_reproject_pixels);
// backward: reprojection -> (dK, d calib pose) -> both se(3) backward passes + the intrinsics backward (which also takes
//           the gradient of K^-1 coming from the rays).  The unfused entry points remain for the module-level API.
__global__ void camera_fwd_k(const float* wfx, const float* wfy, const float* wux,
                             const float* wuy, const float* w_pose,
                             const float* w_calib, const float* wpts, int n, int P, float H,
                             float W, float* K, float* Kinv, float* pose,
                             float* calib, float* pix) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  intrinsics_fwd_dev(i, wfx, wfy, wux, wuy, n, H, W, K, Kinv);
  se3_fwd_dev(i, w_pose, n, pose);
  se3_fwd_dev(i, w_calib, n, calib);
  reproject_fwd_dev(i, wpts, K, calib, n, P, pix);        // reads what this very thread just wrote
}

__global__ void camera_bwd_k(const float* wfx, const float* wfy, const float* wux,
                             const float* wuy, const float* w_pose,
                             const float* w_calib, const float* wpts,
                             const float* K, const float* calib, int n, int P, float H, float W,
                             const float* g_Kinv, const float* g_pose,
                             const float* g_pix, float* tmp_gK, float* tmp_gcalib,
                             float* g_fx, float* g_fy, float* g_ux,
                             float* g_uy, float* g_wpose, float* g_wcalib) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  reproject_bwd_dev(i, wpts, K, calib, g_pix, n, P, tmp_gK, tmp_gcalib);
  se3_bwd_dev(i, w_calib, tmp_gcalib, n, g_wcalib);
  if (g_wpose) se3_bwd_dev(i, w_pose, g_pose, n, g_wpose);
  intrinsics_bwd_dev(i, wfx, wfy, wux, wuy, n, H, W, tmp_gK, g_Kinv, g_fx, g_fy, g_ux, g_uy);
}

}  // namespace

extern "C" int mcnerf_reproject_fwd(const float* wpts, const float* K, const float* Rt, int n_cam, int n_pts,
                                    float* pix, void* stream) {
  MC_ARG(wpts && K && Rt && pix && n_cam > 0 && n_pts > 0);
  reproject_fwd_k<<<cdiv(n_cam, 128), 128, 0, (cudaStream_t)stream>>>(wpts, K, Rt, n_cam, n_pts, pix);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_reproject_bwd(const float* wpts, const float* K, const float* Rt, const float* g_pix, int n_cam,
                                    int n_pts, float* gK, float* gRt, void* stream) {
  MC_ARG(wpts && K && Rt && g_pix && gK && gRt && n_cam > 0 && n_pts > 0);
  reproject_bwd_k<<<cdiv(n_cam, 128), 128, 0, (cudaStream_t)stream>>>(wpts, K, Rt, g_pix, n_cam, n_pts, gK, gRt);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_intrinsics_fwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                                     int n_cam, int img_h, int img_w, float* K, float* Kinv, void* stream) {
  MC_ARG(w_fx && w_fy && w_ux && w_uy && K && n_cam > 0);
  intrinsics_fwd_k<<<cdiv(n_cam, 128), 128, 0, (cudaStream_t)stream>>>(w_fx, w_fy, w_ux, w_uy, n_cam, (float)img_h,
                                                                        (float)img_w, K, Kinv);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_intrinsics_bwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                                     int n_cam, int img_h, int img_w, const float* gK, const float* gKinv,
                                     float* g_fx, float* g_fy, float* g_ux, float* g_uy, void* stream) {
  MC_ARG(w_fx && w_fy && w_ux && w_uy && g_fx && g_fy && g_ux && g_uy && n_cam > 0);
  intrinsics_bwd_k<<<cdiv(n_cam, 128), 128, 0, (cudaStream_t)stream>>>(w_fx, w_fy, w_ux, w_uy, n_cam, (float)img_h,
                                                                        (float)img_w, gK, gKinv, g_fx, g_fy, g_ux, g_uy);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_se3_fwd(const float* wu, int n, float* Rt, void* stream) {
  MC_ARG(wu && Rt && n > 0);
  se3_fwd_k<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(wu, n, Rt);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_se3_bwd(const float* wu, const float* gRt, int n, float* g_wu, void* stream) {
  MC_ARG(wu && gRt && g_wu && n > 0);
  se3_bwd_k<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(wu, gRt, n, g_wu);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_camera_fwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                                 const float* w_pose, const float* w_pose_calib, const float* wpts, int n_cam, int n_pts,
                                 int img_h, int img_w, float* K, float* Kinv, float* pose, float* calib_pose, float* pix,
                                 void* stream) {
  MC_ARG(w_fx && w_fy && w_ux && w_uy && w_pose && w_pose_calib && wpts && K && Kinv && pose && calib_pose && pix &&
         n_cam > 0 && n_pts > 0);
  camera_fwd_k<<<cdiv(n_cam, 64), 64, 0, (cudaStream_t)stream>>>(w_fx, w_fy, w_ux, w_uy, w_pose, w_pose_calib, wpts, n_cam,
                                                                 n_pts, (float)img_h, (float)img_w, K, Kinv, pose,
                                                                 calib_pose, pix);
  MC_LAUNCHED();
  return 0;
}

extern "C" int mcnerf_camera_bwd(const float* w_fx, const float* w_fy, const float* w_ux, const float* w_uy,
                                 const float* w_pose, const float* w_pose_calib, const float* wpts, const float* K,
                                 const float* calib_pose, int n_cam, int n_pts, int img_h, int img_w, const float* g_Kinv,
                                 const float* g_pose, const float* g_pix, float* scratch, float* g_fx, float* g_fy,
                                 float* g_ux, float* g_uy, float* g_w_pose, float* g_w_pose_calib, void* stream) {
  MC_ARG(w_fx && w_fy && w_ux && w_uy && w_pose && w_pose_calib && wpts && K && calib_pose && g_pix && scratch && g_fx &&
         g_fy && g_ux && g_uy && g_w_pose_calib && n_cam > 0 && n_pts > 0 && (g_w_pose == nullptr || g_pose != nullptr));
  camera_bwd_k<<<cdiv(n_cam, 64), 64, 0, (cudaStream_t)stream>>>(
      w_fx, w_fy, w_ux, w_uy, w_pose, w_pose_calib, wpts, K, calib_pose, n_cam, n_pts, (float)img_h, (float)img_w, g_Kinv,
      g_pose, g_pix, scratch, scratch + (size_t)9 * n_cam, g_fx, g_fy, g_ux, g_uy, g_w_pose, g_w_pose_calib);
  MC_LAUNCHED();
  return 0;
}
