// extern "C" doorway to the reference's own CUDA launchers, for cross-checks and
// "kernel to beat" timings on the GPU box.  The launchers themselves are compiled
// UNMODIFIED from /root/reference/pointnet2_tensorflow/tf_ops/{sampling/tf_sampling_g.cu,
// grouping/tf_grouping_g.cu} (see oracle/Makefile); only their prototypes are restated here
// (tf_sampling.cpp:65,94,125,150 and tf_grouping.cpp:66,108,142,173).
// TEST INFRASTRUCTURE ONLY -- never linked into libpcops.so.
#include <cuda_runtime.h>

void probsampleLauncher(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out);
void farthestpointsamplingLauncher(int b, int n, int m, const float *inp, float *temp, int *out);
void gatherpointLauncher(int b, int n, int m, const float *inp, const int *idx, float *out);
void scatteraddpointLauncher(int b, int n, int m, const float *out_g, const int *idx, float *inp_g);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                            int *idx, int *pts_cnt);
void selectionSortLauncher(int b, int n, int m, int k, const float *dist, int *outi, float *out);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out);
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                            float *grad_points);

// All reference launches go to the legacy default stream; callers synchronise the device
// (torch.cuda.synchronize) before and after.
extern "C" {
int ref_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
  probsampleLauncher(b, n, m, inp_p, inp_r, temp, out);
  return (int)cudaGetLastError();
}
int ref_fps(int b, int n, int m, const float *inp, float *temp32n, int *out) {
  farthestpointsamplingLauncher(b, n, m, inp, temp32n, out);
  return (int)cudaGetLastError();
}
int ref_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
  gatherpointLauncher(b, n, m, inp, idx, out);
  return (int)cudaGetLastError();
}
int ref_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
  cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * 3, 0);  // tf_sampling.cpp:174
  scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
  return (int)cudaGetLastError();
}
int ref_query_ball(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                   int *pts_cnt) {
  queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
  return (int)cudaGetLastError();
}
int ref_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
  selectionSortLauncher(b, n, m, k, dist, outi, out);
  return (int)cudaGetLastError();
}
int ref_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
  groupPointLauncher(b, n, c, m, nsample, points, idx, out);
  return (int)cudaGetLastError();
}
int ref_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                         float *grad_points) {
  cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, 0);  // tf_grouping.cpp:204
  groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
  return (int)cudaGetLastError();
}
int ref_sync(void) { return (int)cudaDeviceSynchronize(); }
}
