// salp_capi.cu -- the C ABI of include/salp_b200.h: handle lifetime, device memory, the
// host-buffer entry points.  No torch, no C++ types across the boundary, nothing throws.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "salp_common.cuh"

#define SALP_HOST_RANGES 8           // salp_step_host, K-sorted: env ranges whose D2H overlaps the next range's kernel
#define SALP_HOST_RANGE_MIN 65536    // ... each at least this many envs (the K-sort balances warps WITHIN a range)

struct SalpSim {
  SalpParams params;
  SalpView view;
  SalpScratch scratch;
  int device;
  int obs_dim;
  int64_t launches;
  const char* last_kernel;
  std::string error;
  // device staging for the *_host entry points
  float* d_actions;
  float* d_obs;
  float* d_reward;
  uint8_t* d_terminated;
  uint8_t* d_truncated;
  float* d_terminal_obs;
  double* d_terms;
  int32_t* d_substeps;
  double* d_metrics;
  uint8_t* d_mask;
  float* d_pool_targets;
  float* d_pool_obstacles;
  cudaStream_t host_stream;
  // chunked host step (K-sorted, large batches): copy streams and per-range events
  cudaStream_t h2d_stream, d2h_stream, step_stream2;
  SalpScratch scratch2;        // K-sort scratch of the second compute stream
  cudaEvent_t ev_in[SALP_HOST_RANGES], ev_done[SALP_HOST_RANGES];
  bool chunk_ready;
};

static thread_local std::string g_create_error;

static int fail(SalpSim* h, int code, const std::string& msg) {
  if (h) h->error = msg; else g_create_error = msg;
  return code;
}
static int cuda_fail(SalpSim* h, cudaError_t e, const char* what) {
  return fail(h, SALP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(h, call)                                            \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);   \
  } while (0)

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// least-squares quadratic through the reference's 4 data points (geometry.py:6-10, 17-21);
// a Python host overwrites the result with np.polyfit's own bits (params.py).
static void fit_quadratic(const double* x, const double* y, int n, double out[3]) {
  long double S[5] = {0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    long double xi = x[i], p = 1;
    for (int k = 0; k < 5; k++) { S[k] += p; if (k < 3) T[k] += p * y[i]; p *= xi; }
  }
  // normal equations [S4 S3 S2; S3 S2 S1; S2 S1 S0] [a b c]^T = [T2 T1 T0]^T, Cramer's rule
  long double A[3][3] = {{S[4], S[3], S[2]}, {S[3], S[2], S[1]}, {S[2], S[1], S[0]}};
  long double B[3] = {T[2], T[1], T[0]};
  auto det3 = [](long double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  };
  long double d = det3(A);
  for (int c = 0; c < 3; c++) {
    long double M[3][3];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) M[r][k] = (k == c) ? B[r] : A[r][k];
    out[c] = (double)(det3(M) / d);
  }
}

extern "C" {

int salp_default_params(SalpParams* p) {
  if (!p) return SALP_ERR_INVALID;
  memset(p, 0, sizeof *p);
  const double pi = 3.14159265358979323846;
  p->nozzle_length1 = p->nozzle_length2 = p->nozzle_length3 = 0.05;
  p->nozzle_area = 0.00016;
  p->nozzle_mass = 1.0;
  p->nozzle_gamma = pi / 4;
  p->nozzle_angle_speed = 31 * pi / 30;
  p->dry_mass = 1.0;
  p->init_length = 0.3;
  p->init_width = 0.15;
  p->max_contraction = 0.06;
  p->density = 1000.0;
  p->dt = 0.01;
  p->buoy_mass = 0.195;
  p->skin_mass = 0.145;
  p->tube_mass = 0.414;
  p->tube_volume = pi * ((0.058 / 2) * (0.058 / 2)) * 0.15;
  p->discharge_coefficient = 0.3;
  p->drag_force_ratio = 0.25;
  p->drag_torque_ratio = 0.1;
  const double amf[3] = {0.5, 0.6, 0.6}, amrf[3] = {0.2, 0.2, 0.2}, amt[3] = {0.3, 0.6, 0.6};
  const double tdr[6] = {1.5, 2.5, 2.5, 1.5, 2.5, 1.5}, rdr[6] = {0.1, 0.3, 0.5, 0.2, 0.5, 0.2};
  for (int i = 0; i < 3; i++) {
    p->added_mass_force[i] = amf[i];
    p->added_mass_rate_force[i] = amrf[i];
    p->added_mass_torque[i] = amt[i];
    p->added_mass_rate_torque[i] = amrf[i];
  }
  for (int i = 0; i < 6; i++) { p->trans_drag_range[i] = tdr[i]; p->rot_drag_range[i] = rdr[i]; }
  const double cx[4] = {0.01, 0.02, 0.03, 0.04};
  const double refill[4] = {0.4, 1.0, 1.8, 2.2}, jet[4] = {0.1, 0.3, 0.4, 0.5};
  fit_quadratic(cx, refill, 4, p->refill_poly);
  fit_quadratic(cx, jet, 4, p->jet_poly);
  const double scale = 200.0, margin = 50, width = 900, height = 700;
  p->tank_x_min = (-width / 2 + margin) / scale;
  p->tank_x_max = (width / 2 - margin) / scale;
  p->tank_y_min = (-height / 2 + margin) / scale;
  p->tank_y_max = (height / 2 - margin) / scale;
  p->target_radius = 0.2;
  p->obstacle_radius = 0.2;
  p->out_of_bounds_distance = 5.0;
  p->success_bonus = 500.0;
  p->out_of_bounds_penalty = 200.0;
  p->collision_penalty = 200.0;
  p->timeout_penalty = 50.0;
  p->max_cycles = 500;
  p->num_obstacles = 2;
  p->precision = SALP_PRECISION_MIXED;
  p->randomization = 0;
  return SALP_OK;
}

const char* salp_build_info(void) {
  static char buf[160];
  snprintf(buf, sizeof buf, "salp_b200 abi %d, sm_100a, nvcc %d.%d.%d, built %s", SALP_ABI_VERSION,
           __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, __CUDACC_VER_BUILD__, __DATE__);
  return buf;
}

const char* salp_last_error(salp_handle h) { return h ? h->error.c_str() : g_create_error.c_str(); }
int64_t salp_num_envs(salp_handle h) { return h ? h->view.n : 0; }
int32_t salp_obs_dim(salp_handle h) { return h ? h->obs_dim : 0; }
int64_t salp_launch_count(salp_handle h) { return h ? h->launches : 0; }
const char* salp_last_step_kernel(salp_handle h) { return h ? h->last_kernel : ""; }
int32_t salp_abi_version(void) { return SALP_ABI_VERSION; }
int64_t salp_sizeof_params(void) { return (int64_t)sizeof(SalpParams); }
int64_t salp_sizeof_step_io(void) { return (int64_t)sizeof(SalpStepIO); }

int salp_destroy(salp_handle h) {
  if (!h) return SALP_OK;
  DeviceGuard g(h->device);
  void* ptrs[] = {h->view.f64, h->view.f32, h->view.i32, h->view.status, (void*)h->view.time_table,
                  h->scratch.K, h->scratch.order, h->scratch.hist, h->d_actions, h->d_obs, h->d_reward,
                  h->d_terminated, h->d_truncated, h->d_terminal_obs, h->d_terms, h->d_substeps, h->d_metrics,
                  h->d_mask, h->d_pool_targets, h->d_pool_obstacles};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  if (h->chunk_ready) {
    cudaStreamDestroy(h->h2d_stream);
    cudaStreamDestroy(h->d2h_stream);
    cudaStreamDestroy(h->step_stream2);
    cudaFree(h->scratch2.K); cudaFree(h->scratch2.order); cudaFree(h->scratch2.hist);
    for (int r = 0; r < SALP_HOST_RANGES; r++) { cudaEventDestroy(h->ev_in[r]); cudaEventDestroy(h->ev_done[r]); }
  }
  delete h;
  return SALP_OK;
}

int salp_create(const SalpParams* params, int64_t num_envs, int device, uint64_t seed, int64_t env_id_offset,
                salp_handle* out) {
  if (!out) return fail(nullptr, SALP_ERR_INVALID, "salp_create: out is NULL");
  *out = nullptr;
  if (!params || num_envs <= 0 || num_envs > 0x7fffffff)
    return fail(nullptr, SALP_ERR_INVALID, "salp_create: params NULL or num_envs out of range");
  if (params->num_obstacles < 0 || params->num_obstacles > SALP_MAX_OBSTACLES)
    return fail(nullptr, SALP_ERR_INVALID, "salp_create: num_obstacles out of range");
  if (params->precision != SALP_PRECISION_F64 && params->precision != SALP_PRECISION_MIXED)
    return fail(nullptr, SALP_ERR_INVALID, "salp_create: unknown precision");
  if (!(params->dt > 0)) return fail(nullptr, SALP_ERR_INVALID, "salp_create: dt must be > 0");
  if (params->randomization != 0 && params->precision != SALP_PRECISION_MIXED)
    return fail(nullptr, SALP_ERR_INVALID, "salp_create: randomization needs SALP_PRECISION_MIXED "
                                           "(the float64 reference mode is the deterministic restatement)");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count)
    return fail(nullptr, SALP_ERR_NO_DEVICE,
                std::string("salp_create: no usable CUDA device (this library has no CPU fallback): ") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, SALP_ERR_NO_DEVICE, "salp_create: device is not sm_100 or newer (kernels are built for sm_100a only)");
  DeviceGuard g(device);
  if (!g.ok) return fail(nullptr, SALP_ERR_CUDA, "salp_create: cudaSetDevice failed");

  SalpSim* h = new (std::nothrow) SalpSim();
  if (!h) return fail(nullptr, SALP_ERR_ALLOC, "salp_create: out of host memory");
  memset(&h->view, 0, sizeof h->view);
  memset(&h->scratch, 0, sizeof h->scratch);
  h->d_actions = h->d_obs = h->d_reward = h->d_terminal_obs = nullptr;
  h->d_terminated = h->d_truncated = h->d_mask = nullptr;
  h->d_terms = h->d_metrics = nullptr;
  h->d_substeps = nullptr;
  h->d_pool_targets = h->d_pool_obstacles = nullptr;
  h->host_stream = nullptr;
  h->chunk_ready = false;
  h->params = *params;
  h->device = device;
  h->obs_dim = SALP_OBS_BASE + 2 * params->num_obstacles;
  h->launches = 0;
  h->last_kernel = "";
  const int64_t n = num_envs;
  const int D = h->obs_dim;
  SalpView& v = h->view;
  v.n = n;
  v.env_id_offset = env_id_offset;
  v.seed = seed;
  v.sm_count = prop.multiProcessorCount;
  double* table = nullptr;
#define ALLOC(ptr, bytes)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = cudaMalloc((void**)&(ptr), (size_t)(bytes));                          \
    if (e__ != cudaSuccess) {                                                               \
      int rc = cuda_fail(nullptr, e__, "salp_create: cudaMalloc " #ptr);                    \
      salp_destroy(h);                                                                      \
      return rc == SALP_ERR_CUDA ? SALP_ERR_ALLOC : rc;                                     \
    }                                                                                       \
  } while (0)
  ALLOC(v.f64, sizeof(double) * SALP_NUM_F64_FIELDS * n);
  ALLOC(v.f32, sizeof(float) * SALP_NUM_F32 * n);
  ALLOC(v.i32, sizeof(int32_t) * SALP_NUM_I32 * n);
  ALLOC(v.status, sizeof(int32_t));
  ALLOC(table, sizeof(double) * (SALP_MAX_SUBSTEPS + 1));
  v.time_table = table;
  ALLOC(h->scratch.K, sizeof(int32_t) * n);
  ALLOC(h->scratch.order, sizeof(int32_t) * n);
  ALLOC(h->scratch.hist, sizeof(int32_t) * SALP_SORT_BINS);
  ALLOC(h->d_actions, sizeof(float) * 3 * n);
  ALLOC(h->d_obs, sizeof(float) * D * n);
  ALLOC(h->d_reward, sizeof(float) * n);
  ALLOC(h->d_terminated, n);
  ALLOC(h->d_truncated, n);
  ALLOC(h->d_terminal_obs, sizeof(float) * D * n);
  ALLOC(h->d_terms, sizeof(double) * SALP_NUM_REWARD_TERMS * n);
  ALLOC(h->d_substeps, sizeof(int32_t) * n);
  ALLOC(h->d_metrics, sizeof(double) * SALP_NUM_EPISODE_METRICS * n);
  ALLOC(h->d_mask, n);
#undef ALLOC
  // t_k: k-fold repeated `cycle_time += dt` from 0.0 (robot.py:587, 674) -- NOT k*dt
  {
    std::vector<double> t(SALP_MAX_SUBSTEPS + 1);
    volatile double acc = 0.0;
    for (int k = 0; k <= SALP_MAX_SUBSTEPS; k++) { t[k] = acc; acc = acc + params->dt; }
    e = cudaMemcpy(table, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(v.status, 0, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { int rc = cuda_fail(nullptr, e, "salp_create: table upload"); salp_destroy(h); return rc; }
  }
  int rc = salp_launch_init(h->params, v, h->host_stream);
  if (rc < 0 || (e = cudaStreamSynchronize(h->host_stream)) != cudaSuccess) {
    rc = cuda_fail(nullptr, rc < 0 ? cudaGetLastError() : e, "salp_create: init kernel (is this a sm_100 GPU?)");
    salp_destroy(h);
    return rc;
  }
  h->launches += 1;
  *out = h;
  return SALP_OK;
}

int salp_reset(salp_handle h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!h) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  int rc = salp_launch_reset(h->params, h->view, mask_dev, obs_dev, (cudaStream_t)stream);
  if (rc < 0) return cuda_fail(h, cudaGetLastError(), "salp_reset launch");
  h->launches += rc;
  return SALP_OK;
}

int salp_step(salp_handle h, const SalpStepIO* io, uint32_t flags, void* stream) {
  if (!h || !io) return SALP_ERR_INVALID;
  if (!io->actions || !io->obs || !io->reward || !io->terminated || !io->truncated)
    return fail(h, SALP_ERR_INVALID, "salp_step: actions, obs, reward, terminated and truncated are required");
  DeviceGuard g(h->device);
  int rc = salp_launch_step(h->params, h->view, *io, flags, h->scratch, (cudaStream_t)stream, &h->last_kernel);
  if (rc < 0) return cuda_fail(h, cudaGetLastError(), "salp_step launch");
  h->launches += rc;
  return SALP_OK;
}

int salp_reset_host(salp_handle h, const uint8_t* mask_host, float* obs_host) {
  if (!h) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  cudaStream_t s = h->host_stream;
  const int64_t n = h->view.n;
  if (mask_host) CU(h, cudaMemcpyAsync(h->d_mask, mask_host, n, cudaMemcpyHostToDevice, s));
  int rc = salp_launch_reset(h->params, h->view, mask_host ? h->d_mask : nullptr, h->d_obs, s);
  if (rc < 0) return cuda_fail(h, cudaGetLastError(), "salp_reset_host launch");
  h->launches += rc;
  if (obs_host) CU(h, cudaMemcpyAsync(obs_host, h->d_obs, sizeof(float) * h->obs_dim * n, cudaMemcpyDeviceToHost, s));
  CU(h, cudaStreamSynchronize(s));
  return SALP_OK;
}

// Device-side alias of a caller's HOST buffer, if it has one: page-locked memory (cudaHostAlloc /
// cudaHostRegister, torch's pin_memory) is mapped into the device address space under unified
// addressing.  Pageable memory (a plain numpy array) has none -> nullptr, the staged path is used.
static void* mapped_alias(const void* host_ptr) {
  if (!host_ptr) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return at.devicePointer;
}
static bool zero_copy_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SALP_ZERO_COPY"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

// Host-buffer step.  Two transports, chosen per call from the caller's pointers:
//  * zero-copy (page-locked caller buffers, MIXED precision): the step kernel stores obs / reward /
//    flags / terminal obs straight into the caller's memory as each warp finishes -- the transfer
//    overlaps the integration of the other warps and no copy is queued behind the kernel.  The
//    kernel only ever WRITES those arrays (observation rows are assembled in shared memory).
//    Actions are read in place for batches of at most one warp per SM sub-partition and staged by
//    the copy engine above that.  Used for steps in natural env order only, whichever MIXED step
//    kernel the launcher picks (fused or pipeline: all assemble rows in shared memory).
//  * staged (pageable caller buffers, K-sorted steps, F64): H2D of the actions, kernel on the
//    handle's own device buffers, one D2H per output.
// The optional extras (reward_terms, substeps, episode_metrics) are always staged.
static bool page_locked(const void* p) {
  cudaPointerAttributes at;
  if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// K-sorted host step of a large batch in contiguous env RANGES: each range is planned, K-sorted and
// stepped on its own (sub-view of the state, global env ids kept), and as soon as a range's kernel is
// done its slices of obs / reward / flags / terminal obs go back over PCIe on a second stream while
// the next range integrates; the action slices are uploaded ahead on a third.  Same results as the
// one-launch step (per-env results do not depend on which envs share a warp).  Needs page-locked
// caller buffers (asynchronous copies); otherwise the plain staged path below is used.
static int step_host_ranges(SalpSim* h, const SalpStepIO* io, uint32_t flags, int ranges) {
  const int64_t n = h->view.n;
  const int D = h->obs_dim;
  cudaStream_t s = h->host_stream;
  if (!h->chunk_ready) {
    CU(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CU(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    CU(h, cudaStreamCreateWithFlags(&h->step_stream2, cudaStreamNonBlocking));
    CU(h, cudaMalloc((void**)&h->scratch2.K, sizeof(int32_t) * n));
    CU(h, cudaMalloc((void**)&h->scratch2.order, sizeof(int32_t) * n));
    CU(h, cudaMalloc((void**)&h->scratch2.hist, sizeof(int32_t) * SALP_SORT_BINS));
    for (int r = 0; r < SALP_HOST_RANGES; r++) {
      CU(h, cudaEventCreateWithFlags(&h->ev_in[r], cudaEventDisableTiming));
      CU(h, cudaEventCreateWithFlags(&h->ev_done[r], cudaEventDisableTiming));
    }
    h->chunk_ready = true;
  }
  // Range sizes: equal, except that with six or more ranges the last two are a half and a quarter --
  // what stays exposed at the end of the step is the LAST range's result copy (its kernel shares the
  // GPU with the range before it, on the other compute stream).  Measured e2e / device-timed: 0.888
  // vs 0.868 at 1 M envs (8 ranges); with 4 ranges (262 144 envs) the small tail ranges cost more
  // than they save (0.814 vs 0.833).  SALP_HOST_RANGE_TAIL=0: equal ranges.
  static const bool taper = [] { const char* e = getenv("SALP_HOST_RANGE_TAIL"); return !e || atoi(e) != 0; }();
  int64_t first_of[SALP_HOST_RANGES + 1];
  {
    double w[SALP_HOST_RANGES], total = 0.0;
    for (int r = 0; r < ranges; r++) {
      w[r] = (taper && ranges >= 6 && r == ranges - 1) ? 0.25 : (taper && ranges >= 6 && r == ranges - 2) ? 0.5 : 1.0;
      total += w[r];
    }
    double acc = 0.0;
    first_of[0] = 0;
    for (int r = 0; r < ranges; r++) {
      acc += w[r];
      int64_t e = (int64_t)((double)n * acc / total + 31.0) / 32 * 32;
      first_of[r + 1] = (r == ranges - 1 || e > n) ? n : e;
    }
  }
  for (int r = 0; r < ranges; r++) {
    const int64_t first = first_of[r];
    const int64_t cnt = first_of[r + 1] - first;
    if (cnt <= 0) continue;
    CU(h, cudaMemcpyAsync(h->d_actions + 3 * first, io->actions + 3 * first, sizeof(float) * 3 * cnt,
                          cudaMemcpyHostToDevice, h->h2d_stream));
    CU(h, cudaEventRecord(h->ev_in[r], h->h2d_stream));
  }
  for (int r = 0; r < ranges; r++) {
    const int64_t first = first_of[r];
    const int64_t cnt = first_of[r + 1] - first;
    if (cnt <= 0) continue;
    SalpView v = h->view;
#if SALP_STATE_AOS
    v.f64 += first * SALP_NUM_F64_FIELDS;
    v.f32 += first * SALP_NUM_F32;
    v.i32 += first * SALP_NUM_I32;
#else
#error "step_host_ranges needs the record-per-env state layout"
#endif
    v.n = cnt;
    v.env_id_offset += first;
    if (v.pool_P > 0) {
      v.pool_targets += first * v.pool_P * 2;
      v.pool_obstacles += first * v.pool_P * h->params.num_obstacles * 2;
    }
    SalpStepIO d;
    d.actions = h->d_actions + 3 * first;
    d.obs = h->d_obs + D * first;
    d.reward = h->d_reward + first;
    d.terminated = h->d_terminated + first;
    d.truncated = h->d_truncated + first;
    d.terminal_obs = io->terminal_obs ? h->d_terminal_obs + D * first : nullptr;
    d.reward_terms = io->reward_terms ? h->d_terms + SALP_NUM_REWARD_TERMS * first : nullptr;
    d.substeps = io->substeps ? h->d_substeps + first : nullptr;
    d.episode_metrics = io->episode_metrics ? h->d_metrics + SALP_NUM_EPISODE_METRICS * first : nullptr;
    // two compute streams, alternating: the thin tail of one range's kernel (its longest warps)
    // overlaps the head of the next range's
    cudaStream_t cs = (r & 1) ? h->step_stream2 : s;
    CU(h, cudaStreamWaitEvent(cs, h->ev_in[r], 0));
    int rc = salp_launch_step(h->params, v, d, flags, (r & 1) ? h->scratch2 : h->scratch, cs, &h->last_kernel);
    if (rc < 0) return cuda_fail(h, cudaGetLastError(), "salp_step_host launch (range)");
    h->launches += rc;
    CU(h, cudaEventRecord(h->ev_done[r], cs));
    cudaStream_t c = h->d2h_stream;
    CU(h, cudaStreamWaitEvent(c, h->ev_done[r], 0));
    CU(h, cudaMemcpyAsync(io->obs + D * first, d.obs, sizeof(float) * D * cnt, cudaMemcpyDeviceToHost, c));
    CU(h, cudaMemcpyAsync(io->reward + first, d.reward, sizeof(float) * cnt, cudaMemcpyDeviceToHost, c));
    CU(h, cudaMemcpyAsync(io->terminated + first, d.terminated, cnt, cudaMemcpyDeviceToHost, c));
    CU(h, cudaMemcpyAsync(io->truncated + first, d.truncated, cnt, cudaMemcpyDeviceToHost, c));
    if (io->terminal_obs)
      CU(h, cudaMemcpyAsync(io->terminal_obs + D * first, d.terminal_obs, sizeof(float) * D * cnt, cudaMemcpyDeviceToHost, c));
    if (io->reward_terms)
      CU(h, cudaMemcpyAsync(io->reward_terms + SALP_NUM_REWARD_TERMS * first, d.reward_terms,
                            sizeof(double) * SALP_NUM_REWARD_TERMS * cnt, cudaMemcpyDeviceToHost, c));
    if (io->substeps)
      CU(h, cudaMemcpyAsync(io->substeps + first, d.substeps, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, c));
    if (io->episode_metrics)
      CU(h, cudaMemcpyAsync(io->episode_metrics + SALP_NUM_EPISODE_METRICS * first, d.episode_metrics,
                            sizeof(double) * SALP_NUM_EPISODE_METRICS * cnt, cudaMemcpyDeviceToHost, c));
  }
  CU(h, cudaStreamSynchronize(h->d2h_stream));     // (waits for every range's kernel through the events)
  CU(h, cudaStreamSynchronize(s));
  CU(h, cudaStreamSynchronize(h->step_stream2));
  return SALP_OK;
}

int salp_step_host(salp_handle h, const SalpStepIO* io, uint32_t flags) {
  if (!h || !io) return SALP_ERR_INVALID;
  if (!io->actions || !io->obs || !io->reward || !io->terminated || !io->truncated)
    return fail(h, SALP_ERR_INVALID, "salp_step_host: actions, obs, reward, terminated and truncated are required");
  DeviceGuard g(h->device);
  cudaStream_t s = h->host_stream;
  const int64_t n = h->view.n;
  const int D = h->obs_dim;
  if ((flags & SALP_STEP_SORT_BY_K) && n >= 2 * (int64_t)SALP_HOST_RANGE_MIN && page_locked(io->obs) &&
      page_locked(io->actions) && page_locked(io->reward) && page_locked(io->terminated) && page_locked(io->truncated)) {
    int ranges = (int)(n / SALP_HOST_RANGE_MIN);
    ranges = ranges > SALP_HOST_RANGES ? SALP_HOST_RANGES : ranges;
    return step_host_ranges(h, io, flags, ranges);
  }
  // (K-sorted steps visit the envs in scattered order: 40-byte rows make poor PCIe writes --
  //  measured 12.7 ms vs 8.2 ms staged at 1 M envs -- so they keep the staged transport)
  const bool zc_ok = zero_copy_enabled() && h->params.precision == SALP_PRECISION_MIXED &&
                     !(flags & SALP_STEP_SORT_BY_K);
  float* z_obs = zc_ok ? (float*)mapped_alias(io->obs) : nullptr;
  float* z_reward = zc_ok ? (float*)mapped_alias(io->reward) : nullptr;
  uint8_t* z_term = zc_ok ? (uint8_t*)mapped_alias(io->terminated) : nullptr;
  uint8_t* z_trunc = zc_ok ? (uint8_t*)mapped_alias(io->truncated) : nullptr;
  float* z_tobs = zc_ok ? (float*)mapped_alias(io->terminal_obs) : nullptr;
  const float* z_act = (zc_ok && n <= (int64_t)h->view.sm_count * 4 * 32)
                           ? (const float*)mapped_alias(io->actions) : nullptr;
  if (!z_act) CU(h, cudaMemcpyAsync(h->d_actions, io->actions, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
  SalpStepIO d;
  d.actions = z_act ? z_act : h->d_actions;
  d.obs = z_obs ? z_obs : h->d_obs;
  d.reward = z_reward ? z_reward : h->d_reward;
  d.terminated = z_term ? z_term : h->d_terminated;
  d.truncated = z_trunc ? z_trunc : h->d_truncated;
  d.terminal_obs = io->terminal_obs ? (z_tobs ? z_tobs : h->d_terminal_obs) : nullptr;
  d.reward_terms = io->reward_terms ? h->d_terms : nullptr;
  d.substeps = io->substeps ? h->d_substeps : nullptr;
  d.episode_metrics = io->episode_metrics ? h->d_metrics : nullptr;
  int rc = salp_launch_step(h->params, h->view, d, flags, h->scratch, s, &h->last_kernel);
  if (rc < 0) return cuda_fail(h, cudaGetLastError(), "salp_step_host launch");
  h->launches += rc;
  if (!z_obs) CU(h, cudaMemcpyAsync(io->obs, d.obs, sizeof(float) * D * n, cudaMemcpyDeviceToHost, s));
  if (!z_reward) CU(h, cudaMemcpyAsync(io->reward, d.reward, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  if (!z_term) CU(h, cudaMemcpyAsync(io->terminated, d.terminated, n, cudaMemcpyDeviceToHost, s));
  if (!z_trunc) CU(h, cudaMemcpyAsync(io->truncated, d.truncated, n, cudaMemcpyDeviceToHost, s));
  if (io->terminal_obs && !z_tobs)
    CU(h, cudaMemcpyAsync(io->terminal_obs, d.terminal_obs, sizeof(float) * D * n, cudaMemcpyDeviceToHost, s));
  if (io->reward_terms)
    CU(h, cudaMemcpyAsync(io->reward_terms, d.reward_terms, sizeof(double) * SALP_NUM_REWARD_TERMS * n,
                          cudaMemcpyDeviceToHost, s));
  if (io->substeps)
    CU(h, cudaMemcpyAsync(io->substeps, d.substeps, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  if (io->episode_metrics)
    CU(h, cudaMemcpyAsync(io->episode_metrics, d.episode_metrics, sizeof(double) * SALP_NUM_EPISODE_METRICS * n,
                          cudaMemcpyDeviceToHost, s));
  CU(h, cudaStreamSynchronize(s));
  return SALP_OK;
}

int salp_set_scene_pool(salp_handle h, const float* targets_host, const float* obstacles_host,
                        int64_t scenes_per_env) {
  if (!h) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  CU(h, cudaDeviceSynchronize());
  if (h->d_pool_targets) cudaFree(h->d_pool_targets);
  if (h->d_pool_obstacles) cudaFree(h->d_pool_obstacles);
  h->d_pool_targets = h->d_pool_obstacles = nullptr;
  h->view.pool_targets = h->view.pool_obstacles = nullptr;
  h->view.pool_P = 0;
  if (scenes_per_env <= 0) return SALP_OK;
  if (!targets_host || (h->params.num_obstacles > 0 && !obstacles_host))
    return fail(h, SALP_ERR_INVALID, "salp_set_scene_pool: NULL scene arrays");
  const int64_t n = h->view.n, P = scenes_per_env;
  size_t nt = sizeof(float) * (size_t)n * P * 2;
  size_t no = sizeof(float) * (size_t)n * P * h->params.num_obstacles * 2;
  CU(h, cudaMalloc((void**)&h->d_pool_targets, nt));
  CU(h, cudaMalloc((void**)&h->d_pool_obstacles, no ? no : 4));
  CU(h, cudaMemcpy(h->d_pool_targets, targets_host, nt, cudaMemcpyHostToDevice));
  if (no) CU(h, cudaMemcpy(h->d_pool_obstacles, obstacles_host, no, cudaMemcpyHostToDevice));
  h->view.pool_targets = h->d_pool_targets;
  h->view.pool_obstacles = h->d_pool_obstacles;
  h->view.pool_P = P;
  return SALP_OK;
}

// address of env 0's element of `field`, element size and byte stride between consecutive envs
static int column_of(SalpSim* h, int32_t field, char** base, size_t* elem, size_t* stride) {
  const int64_t n = h->view.n;
  int64_t local, nfields;
  char* arr;
  if (field >= 0 && field < SALP_NUM_F64_FIELDS) {
    local = field; nfields = SALP_NUM_F64_FIELDS; arr = (char*)h->view.f64; *elem = sizeof(double);
  } else if (field >= SALP_F32_BASE && field < SALP_F32_END) {
    local = field - SALP_F32_BASE; nfields = SALP_NUM_F32; arr = (char*)h->view.f32; *elem = sizeof(float);
  } else if (field >= SALP_I32_BASE && field < SALP_I32_END) {
    local = field - SALP_I32_BASE; nfields = SALP_NUM_I32; arr = (char*)h->view.i32; *elem = sizeof(int32_t);
  } else {
    return fail(h, SALP_ERR_INVALID, "unknown SalpField id");
  }
#if SALP_STATE_AOS
  (void)n;
  *base = arr + *elem * local;
  *stride = *elem * nfields;
#else
  (void)nfields;
  *base = arr + *elem * local * n;
  *stride = *elem;
#endif
  return SALP_OK;
}

int salp_get_state(salp_handle h, int32_t field, void* host_dst, int64_t first, int64_t count) {
  if (!h || !host_dst || first < 0 || count < 0 || first + count > h->view.n) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  char* base; size_t elem, stride;
  int rc = column_of(h, field, &base, &elem, &stride);
  if (rc) return rc;
  CU(h, cudaDeviceSynchronize());
  if (count) CU(h, cudaMemcpy2D(host_dst, elem, base + stride * first, stride, elem, (size_t)count, cudaMemcpyDeviceToHost));
  return SALP_OK;
}

int salp_set_state(salp_handle h, int32_t field, const void* host_src, int64_t first, int64_t count) {
  if (!h || !host_src || first < 0 || count < 0 || first + count > h->view.n) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  char* base; size_t elem, stride;
  int rc = column_of(h, field, &base, &elem, &stride);
  if (rc) return rc;
  CU(h, cudaDeviceSynchronize());
  if (count) CU(h, cudaMemcpy2D(base + stride * first, stride, host_src, elem, elem, (size_t)count, cudaMemcpyHostToDevice));
  return SALP_OK;
}

int salp_state_ptr(salp_handle h, int32_t field, void** dev_ptr, int64_t* stride_bytes) {
  if (!h || !dev_ptr || !stride_bytes) return SALP_ERR_INVALID;
  char* base; size_t elem, stride;
  int rc = column_of(h, field, &base, &elem, &stride);
  if (rc) return rc;
  *dev_ptr = base;
  *stride_bytes = (int64_t)stride;
  return SALP_OK;
}

int salp_trace_cycle(salp_handle h, int64_t env, const float* action_host, double* trace_host, int32_t capacity,
                     int32_t* substeps_out) {
  if (!h || !action_host || !trace_host || !substeps_out || capacity <= 0 || env < 0 || env >= h->view.n)
    return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  if (capacity > SALP_MAX_SUBSTEPS) capacity = SALP_MAX_SUBSTEPS;
  double* d_trace = nullptr;
  int32_t* d_K = nullptr;
  const size_t bytes = sizeof(double) * SALP_TRACE_WIDTH * (size_t)capacity;
  CU(h, cudaMalloc((void**)&d_trace, bytes));
  cudaError_t e = cudaMalloc((void**)&d_K, sizeof(int32_t));
  int rc = SALP_OK;
  if (e != cudaSuccess) rc = cuda_fail(h, e, "salp_trace_cycle: cudaMalloc");
  if (rc == SALP_OK && salp_launch_trace(h->params, h->view, env, action_host, d_trace, capacity, d_K, h->host_stream) < 0)
    rc = cuda_fail(h, cudaGetLastError(), "salp_trace_cycle launch");
  int32_t K = 0;
  if (rc == SALP_OK && (e = cudaMemcpyAsync(&K, d_K, sizeof K, cudaMemcpyDeviceToHost, h->host_stream)) != cudaSuccess)
    rc = cuda_fail(h, e, "salp_trace_cycle: copy K");
  if (rc == SALP_OK && (e = cudaStreamSynchronize(h->host_stream)) != cudaSuccess)
    rc = cuda_fail(h, e, "salp_trace_cycle: kernel");
  if (rc == SALP_OK) {
    const int rows = K < capacity ? K : capacity;
    if (rows > 0 && (e = cudaMemcpy(trace_host, d_trace, sizeof(double) * SALP_TRACE_WIDTH * rows, cudaMemcpyDeviceToHost)) != cudaSuccess)
      rc = cuda_fail(h, e, "salp_trace_cycle: copy trace");
    *substeps_out = K;
    h->launches += 1;
  }
  cudaFree(d_trace);
  if (d_K) cudaFree(d_K);
  return rc;
}

int salp_check(salp_handle h) {
  if (!h) return SALP_ERR_INVALID;
  DeviceGuard g(h->device);
  CU(h, cudaDeviceSynchronize());
  int32_t st = 0;
  CU(h, cudaMemcpy(&st, h->view.status, sizeof st, cudaMemcpyDeviceToHost));
  if (st == SALP_ERR_RANGE) return fail(h, st, "an action drove a cycle past SALP_MAX_SUBSTEPS (non-finite or out-of-Box action)");
  if (st == SALP_ERR_HANDOFF) return fail(h, st, "pipeline kernel: a warp read a ring row that was not written for it (hand-off race)");
  if (st != 0) return fail(h, st, "device-side error");
  return SALP_OK;
}

int salp_probe_fp32_peak(int device, int millis, double* tflops_out) {
  if (!tflops_out || millis <= 0) return SALP_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SALP_ERR_NO_DEVICE;
  DeviceGuard g(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SALP_ERR_CUDA;
  float* scratch = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc((void**)&scratch, 4) != cudaSuccess) return SALP_ERR_ALLOC;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8;      // 2048 resident threads per SM
  const int iters = 4096;
  const double flop_per_launch = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
  double best = 0.0, spent = 0.0;
  int rc = SALP_OK;
  for (int rep = 0; rep < 1000 && spent < millis; rep++) {
    cudaEventRecord(e0, 0);
    if (salp_launch_ffma_probe(scratch, blocks, iters, 0) < 0) { rc = SALP_ERR_CUDA; break; }
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = SALP_ERR_CUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    spent += ms;
    if (rep >= 2 && ms > 0) {                            // sustained: the LAST launches decide
      double tf = flop_per_launch / (ms * 1e-3) / 1e12;
      best = tf;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(scratch);
  *tflops_out = best;
  return rc;
}

}  // extern "C"
