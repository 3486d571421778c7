// Error plumbing, device queries and the small utility kernels (Adam, column mean).
#include "common.cuh"

namespace cfl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local cudaEvent_t g_timer[2] = {nullptr, nullptr};
void timer_record(int which, cudaStream_t st) {
  if (g_timer[0] && g_timer[1]) (void)cudaEventRecord(g_timer[which], st);
}

struct DevProps { int valid; int sms; int major; int minor; };
static DevProps g_props[64];

static int load_props(int* dev_out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return CFL_ERR_DEVICE;
  }
  if (!g_props[dev].valid) {
    int sms = 0, mj = 0, mn = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
      set_error("cudaDeviceGetAttribute failed");
      (void)cudaGetLastError();
      return CFL_ERR_DEVICE;
    }
    g_props[dev].sms = sms; g_props[dev].major = mj; g_props[dev].minor = mn;
    g_props[dev].valid = 1;
  }
  *dev_out = dev;
  return CFL_OK;
}

int device_check() {
  int dev;
  int st = load_props(&dev);
  if (st != CFL_OK) return st;
  if (g_props[dev].major != 10) {
    set_error("device %d is sm_%d%d; this library is sm_100a only (no fallback)", dev,
              g_props[dev].major, g_props[dev].minor);
    return CFL_ERR_DEVICE;
  }
  return CFL_OK;
}

int sm_count() {
  int dev;
  if (load_props(&dev) != CFL_OK) return 0;
  return g_props[dev].sms;
}

// ---- Adam (TF-1.x epsilon placement) ----------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr_t, float b1, float b2, float eps, float gscale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                float* __restrict__ m, float* __restrict__ v, int64_t n,
                                const int* __restrict__ step_dev, float lr, float b1, float b2, float eps,
                                float gscale) {
  const int t = *step_dev;
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ---- column mean (centring vector) ------------------------------------------------------
// stage 1: each block sums a contiguous slab of rows in double, fixed order per block.
__global__ void col_sum_partial(const float* __restrict__ E, int64_t N, int d, int64_t lde,
                                int64_t rows_per_block, double* __restrict__ part) {
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block; if (r1 > N) r1 = N;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    double acc = 0.0;
    for (int64_t r = r0; r < r1; ++r) acc += (double)E[r * lde + j];
    part[(int64_t)blockIdx.x * d + j] = acc;
  }
}
__global__ void col_sum_final(const double* __restrict__ part, int nblocks, int d, int64_t N,
                              float* __restrict__ mu) {
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += part[(int64_t)b * d + j];
    mu[j] = (float)(acc / (double)N);
  }
}

}  // namespace cfl

using namespace cfl;

extern "C" {

const char* cfl_last_error(void) { return g_err; }
int cfl_version(void) { return 100; }

int cfl_device_info(int* sms, int* mj, int* mn) {
  int dev;
  int st = load_props(&dev);
  if (st != CFL_OK) return st;
  if (sms) *sms = g_props[dev].sms;
  if (mj) *mj = g_props[dev].major;
  if (mn) *mn = g_props[dev].minor;
  return CFL_OK;
}

int cfl_set_kernel_timer(void* start_event, void* stop_event) {
  g_timer[0] = (cudaEvent_t)start_event;
  g_timer[1] = (cudaEvent_t)stop_event;
  return CFL_OK;
}

int cfl_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr,
                  float beta1, float beta2, float eps, float grad_scale, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(p && g && m && v && n >= 0 && step >= 1, CFL_ERR_INVALID, "adam: bad arguments");
  if (n == 0) return CFL_OK;
  double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  int blocks = (int)((n + 255) / 256);
  int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr_t, beta1, beta2,
                                                       eps, grad_scale);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

int cfl_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev, float lr,
                      float beta1, float beta2, float eps, float grad_scale, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(p && g && m && v && step_dev && n >= 0, CFL_ERR_INVALID, "adam_dev: bad arguments");
  if (n == 0) return CFL_OK;
  int blocks = (int)((n + 255) / 256);
  int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, step_dev, lr, beta1, beta2, eps,
                                                           grad_scale);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

static int col_mean_blocks(int64_t N) {
  int64_t b = (N + 4095) / 4096;
  if (b < 1) b = 1;
  if (b > 1024) b = 1024;
  return (int)b;
}
size_t cfl_col_mean_workspace_bytes(int64_t N, int d) {
  return align_up((size_t)col_mean_blocks(N) * d * sizeof(double), 256) + 256;
}
int cfl_col_mean(const float* E, int64_t N, int d, int64_t lde, float* mu, void* ws,
                 size_t ws_bytes, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(E && mu && N > 0 && d > 0 && lde >= d, CFL_ERR_INVALID, "col_mean: bad arguments");
  CFL_REQUIRE(ws_bytes >= cfl_col_mean_workspace_bytes(N, d), CFL_ERR_WORKSPACE,
              "col_mean: workspace too small");
  Workspace W(ws, ws_bytes);
  int nb = col_mean_blocks(N);
  double* part = W.take<double>((size_t)nb * d);
  int64_t rpb = (N + nb - 1) / nb;
  col_sum_partial<<<nb, 128, 0, (cudaStream_t)stream>>>(E, N, d, lde, rpb, part);
  CFL_LAUNCH_CHECK();
  col_sum_final<<<1, 128, 0, (cudaStream_t)stream>>>(part, nb, d, N, mu);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

}  // extern "C"
