// Host-boundary packing kernels: what sits between a host batch and the op path on either side.
//
//   pc_unpack_features   the feature prologue of the reference's input pipeline, attention_points/train.py:95-98:
//                        colors = tf.div(tf.cast(colors, tf.float32), 255); features = tf.concat([colors, normals], 2).
//                        The host ships the colours as the uint8 they are stored as (3 bytes per point instead of 12)
//                        and the division happens here -- IEEE fp32 division, the same bits as the TF / numpy form.
//   pc_narrow_indices_u16  int32 -> uint16 for index tensors whose values are known to be < 65536 (FPS / ball / three_nn
//                        indices of clouds of at most 65536 points, pts_cnt <= nsample): halves the device-to-host
//                        bytes of a step's integer result; lossless under that precondition, saturating otherwise.
#include "common.cuh"

namespace pc {
namespace {

// one thread per point: 3 colour bytes + 3 normal floats -> 6 floats.  Byte loads / 4-byte stores: the kernel moves
// 39 bytes per point and runs once per batch (3.5 MB at B = 16), far from any bound that matters.
__global__ void __launch_bounds__(256)
unpack_features_kernel(size_t rows, const unsigned char *__restrict__ colors, const float *__restrict__ normals,
                       float *__restrict__ feat) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    const unsigned char *c = colors + r * 3;
    const float *nrm = normals + r * 3;
    float *f = feat + r * 6;
    f[0] = __fdiv_rn((float)c[0], 255.0f);
    f[1] = __fdiv_rn((float)c[1], 255.0f);
    f[2] = __fdiv_rn((float)c[2], 255.0f);
    f[3] = nrm[0]; f[4] = nrm[1]; f[5] = nrm[2];
  }
}

__global__ void __launch_bounds__(256)
narrow_u16_kernel(size_t count, const int *__restrict__ src, unsigned short *__restrict__ dst, bool vec) {
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  auto sat = [](int v) { return (unsigned)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); };
  if (vec) {  // 8 indices per thread: two 128-bit loads, one 128-bit store
    const size_t n8 = count / 8;
    for (size_t i = t0; i < n8; i += stride) {
      const int4 a = __ldg(reinterpret_cast<const int4 *>(src) + 2 * i), b = __ldg(reinterpret_cast<const int4 *>(src) + 2 * i + 1);
      uint4 o;
      o.x = sat(a.x) | (sat(a.y) << 16); o.y = sat(a.z) | (sat(a.w) << 16);
      o.z = sat(b.x) | (sat(b.y) << 16); o.w = sat(b.z) | (sat(b.w) << 16);
      reinterpret_cast<uint4 *>(dst)[i] = o;
    }
    for (size_t i = n8 * 8 + t0; i < count; i += stride) dst[i] = (unsigned short)sat(src[i]);
  } else {
    for (size_t i = t0; i < count; i += stride) dst[i] = (unsigned short)sat(src[i]);
  }
}

}  // namespace
}  // namespace pc

extern "C" int pc_unpack_features(size_t rows, const unsigned char *colors, const float *normals, float *feat,
                                  pc_stream_t stream) {
  if (rows == 0) return PC_OK;
  if (!colors || !normals || !feat) return PC_ERR_INVALID_ARGUMENT;
  const size_t need = (rows + 255) / 256;
  const int grid = pc::resident_grid((const void *)pc::unpack_features_kernel, 256, 0, need);
  pc::unpack_features_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, colors, normals, feat);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_narrow_indices_u16(size_t count, const int *src, unsigned short *dst, pc_stream_t stream) {
  if (count == 0) return PC_OK;
  if (!src || !dst) return PC_ERR_INVALID_ARGUMENT;
  const bool vec = pc::aligned16(src) && pc::aligned16(dst);
  const size_t need = ((vec ? count / 8 + 8 : count) + 255) / 256;
  const int grid = pc::resident_grid((const void *)pc::narrow_u16_kernel, 256, 0, need);
  pc::narrow_u16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(count, src, dst, vec);
  PC_RETURN_LAUNCH_STATUS();
}
