// Keypoint ingestion and blending on the device: the element-wise data
// preparation either side of the fit, so that detections copied to the GPU once go straight into
// the batch targets.  JSON parsing itself stays host-side text IO (data_parser.py).
//
//   pack_keypoints_kernel   reference data_parser.py:57-104 read_keypoints: raw OpenPose blocks
//                           -> [B,K,3] rows body | left hand | right hand | face[17:68] | face[0:17]
//   keypoint_masks_kernel   reference fit_single_frame.py:276-294 (+ data_parser.py:159-171):
//                           gt / conf split, low-confidence mask, joint weights, trimmed init joints
//   blend_keypoints_kernel  reference keypoints_blending.py:337-369: MMPose confidences moved onto
//                           OpenPose's scale by per-keypoint z-scores, the more confident detection
//                           wins, face rows from OpenPose
// All float32 arithmetic is written without fused multiply-adds so the results equal the numpy
// mirrors (and the reference) bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sfx {

__global__ void pack_keypoints_kernel(const float* __restrict__ body, const float* __restrict__ lhand,
                                      const float* __restrict__ rhand, const float* __restrict__ face,
                                      int B, int nb, int nface, int use_contour, int K,
                                      float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    const int b = idx / K, k = idx % K;
    const float* src;
    if (k < nb) src = body + ((size_t)b * nb + k) * 3;
    else if (k < nb + 21) src = lhand + ((size_t)b * 21 + (k - nb)) * 3;
    else if (k < nb + 42) src = rhand + ((size_t)b * 21 + (k - nb - 21)) * 3;
    else if (k < nb + 42 + 51) src = face + ((size_t)b * nface + 17 + (k - nb - 42)) * 3;
    else src = face + ((size_t)b * nface + (k - nb - 42 - 51)) * 3;
    (void)use_contour;
    for (int c = 0; c < 3; ++c) out[(size_t)idx * 3 + c] = src[c];
}

template <typename T>
__global__ void keypoint_masks_kernel(const float* __restrict__ kp, const float* __restrict__ base_jw,
                                      const int* __restrict__ init_idx, int n_init, int nb, float thr,
                                      int B, int K, T* __restrict__ gt, T* __restrict__ conf,
                                      T* __restrict__ jw, unsigned char* __restrict__ lowconf,
                                      unsigned char* __restrict__ init_mask) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    const int k = idx % K;
    const float x = kp[(size_t)idx * 3], y = kp[(size_t)idx * 3 + 1], c = kp[(size_t)idx * 3 + 2];
    // thresholds = [confidence_threshold] * NUM_BODY_JOINTS + [0] * 42 + [0] * 68
    const bool low = c < (k < nb ? thr : 0.f);
    bool is_init = false;
    for (int i = 0; i < n_init; ++i) is_init = is_init || init_idx[i] == k;
    gt[(size_t)idx * 2] = (T)x;
    gt[(size_t)idx * 2 + 1] = (T)y;
    conf[idx] = (T)c;
    jw[idx] = low ? (T)0 : (T)base_jw[k];
    lowconf[idx] = low ? 1 : 0;
    init_mask[idx] = (is_init && x != 0.f && y != 0.f && !low) ? 1 : 0;
}

// stats: [4][n_pairs] = mmpose_means | mmpose_stds | openpose_means | openpose_stds
__global__ void blend_keypoints_kernel(const float* __restrict__ op, const float* __restrict__ mm,
                                       const float* __restrict__ stats, const int* __restrict__ pair_mm,
                                       const int* __restrict__ pair_op, int n_pairs, int B,
                                       float* __restrict__ out) {
    constexpr int NOP = 135, NMM = 136, FACE0 = 67;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = n_pairs + (NOP - FACE0);
    if (idx >= B * per) return;
    const int b = idx / per, t = idx % per;
    auto clip01 = [](float v) { return fminf(fmaxf(v, 0.f), 1.f); };
    if (t >= n_pairs) {                        // face rows: OpenPose, confidence clipped to [0, 1]
        const int r = FACE0 + (t - n_pairs);
        const float* s = op + ((size_t)b * NOP + r) * 3;
        float* d = out + ((size_t)b * NOP + r) * 3;
        d[0] = s[0];
        d[1] = s[1];
        d[2] = clip01(s[2]);
        return;
    }
    const int mi = pair_mm[t], oi = pair_op[t];
    const float* so = op + ((size_t)b * NOP + oi) * 3;
    const float* sm = mm + ((size_t)b * NMM + mi) * 3;
    const float opc = clip01(so[2]);
    float z = __fdiv_rn(__fsub_rn(sm[2], stats[t]), stats[n_pairs + t]);
    z = clip01(__fadd_rn(__fmul_rn(z, stats[3 * n_pairs + t]), stats[2 * n_pairs + t]));
    const bool take = z > opc;
    float* d = out + ((size_t)b * NOP + oi) * 3;
    d[0] = take ? sm[0] : so[0];
    d[1] = take ? sm[1] : so[1];
    d[2] = take ? z : opc;
}

}  // namespace sfx
