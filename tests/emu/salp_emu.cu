// salp_emu.cu -- TEST INFRASTRUCTURE ONLY.  Not part of the product, never shipped or loaded by it.
//
// Compiles the very same per-env step/reset bodies as the CUDA kernels (csrc/salp_env.cuh, which
// are __host__ __device__) for the HOST and runs them in a plain loop over envs, behind the
// host-buffer subset of the C ABI (salp_create / salp_reset_host / salp_step_host /
// salp_get_state / salp_set_state / salp_set_scene_pool / salp_check).  It exists so that the
// GPU-less CI container (pytest -m "not gpu") can check the kernel LOGIC -- phase plan, K,
// float32/float64 typing quirks, mixed-precision loop structure, reset/auto-reset bookkeeping --
// against the oracle before GPU minutes are spent.  Arithmetic differs from the device in FMA
// contraction and libm, so the GPU parity tests (pytest -m gpu) remain the real gate.
//
// Build: tests/emu/build_emu.py -> tests/emu/_build/libsalp_emu.so  (nvcc, host code only matters)
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../grasp_lab_salp_b200/csrc/salp_env.cuh"

struct SalpSim {
  SalpParams params;
  SalpView view;
  int obs_dim;
  int64_t steps;
  int32_t status;
  std::string error;
  std::vector<double> f64, table;
  std::vector<float> f32, pool_t, pool_o;
  std::vector<int32_t> i32;
};

extern "C" {

const char* salp_build_info(void) { return "salp_emu (host build of the device step body; tests only)"; }
const char* salp_last_error(salp_handle h) { return h ? h->error.c_str() : ""; }
int64_t salp_num_envs(salp_handle h) { return h ? h->view.n : 0; }
int32_t salp_obs_dim(salp_handle h) { return h ? h->obs_dim : 0; }
int64_t salp_launch_count(salp_handle h) { return h ? h->steps : 0; }
int salp_destroy(salp_handle h) { delete h; return SALP_OK; }

int salp_create(const SalpParams* params, int64_t n, int device, uint64_t seed, int64_t env_id_offset,
                salp_handle* out) {
  (void)device;
  if (!params || !out || n <= 0) return SALP_ERR_INVALID;
  SalpSim* h = new SalpSim();
  h->params = *params;
  h->obs_dim = SALP_OBS_BASE + 2 * params->num_obstacles;
  h->steps = 0;
  h->status = 0;
  h->f64.assign((size_t)SALP_NUM_F64_FIELDS * n, 0.0);
  h->f32.assign((size_t)SALP_NUM_F32 * n, 0.f);
  h->i32.assign((size_t)SALP_NUM_I32 * n, 0);
  h->table.resize(SALP_MAX_SUBSTEPS + 1);
  volatile double acc = 0.0;
  for (int k = 0; k <= SALP_MAX_SUBSTEPS; k++) { h->table[k] = acc; acc = acc + params->dt; }
  memset(&h->view, 0, sizeof h->view);
  h->view.f64 = h->f64.data();
  h->view.f32 = h->f32.data();
  h->view.i32 = h->i32.data();
  h->view.n = n;
  h->view.env_id_offset = env_id_offset;
  h->view.seed = seed;
  h->view.status = &h->status;
  h->view.time_table = h->table.data();
  for (int64_t i = 0; i < n; i++) env_init(h->params, h->view, i);
  *out = h;
  return SALP_OK;
}

int salp_reset_host(salp_handle h, const uint8_t* mask, float* obs) {
  if (!h) return SALP_ERR_INVALID;
  for (int64_t i = 0; i < h->view.n; i++)
    if (!mask || mask[i]) env_reset(h->params, h->view, i, obs ? obs + i * h->obs_dim : nullptr);
  return SALP_OK;
}

int salp_step_host(salp_handle h, const SalpStepIO* io, uint32_t flags) {
  if (!h || !io || !io->actions || !io->obs || !io->reward || !io->terminated || !io->truncated)
    return SALP_ERR_INVALID;
  SalpDerived dv = make_derived(h->params);
  if (flags & SALP_STEP_GENERIC) dv.axisym = 0;
  for (int64_t i = 0; i < h->view.n; i++) {
    if (h->params.precision == SALP_PRECISION_F64)
      env_step<SALP_PRECISION_F64>(h->params, dv, h->view, *io, flags, i);
    else if (h->params.randomization != 0)
      env_step<SALP_PRECISION_MIXED_RANDOMIZED>(h->params, dv, h->view, *io, flags, i);
    else
      env_step<SALP_PRECISION_MIXED>(h->params, dv, h->view, *io, flags, i);
  }
  h->steps++;
  return SALP_OK;
}

int salp_set_scene_pool(salp_handle h, const float* t, const float* o, int64_t P) {
  if (!h) return SALP_ERR_INVALID;
  h->view.pool_P = 0;
  h->view.pool_targets = h->view.pool_obstacles = nullptr;
  if (P <= 0) return SALP_OK;
  int64_t n = h->view.n;
  h->pool_t.assign(t, t + n * P * 2);
  h->pool_o.assign(o, o + n * P * h->params.num_obstacles * 2);
  h->view.pool_targets = h->pool_t.data();
  h->view.pool_obstacles = h->pool_o.data();
  h->view.pool_P = P;
  return SALP_OK;
}

static int column_of(SalpSim* h, int32_t field, char** base, size_t* elem, size_t* stride) {
  const int64_t n = h->view.n;
  int64_t local, nfields;
  char* arr;
  if (field >= 0 && field < SALP_NUM_F64_FIELDS) { local = field; nfields = SALP_NUM_F64_FIELDS; arr = (char*)h->view.f64; *elem = 8; }
  else if (field >= SALP_F32_BASE && field < SALP_F32_END) { local = field - SALP_F32_BASE; nfields = SALP_NUM_F32; arr = (char*)h->view.f32; *elem = 4; }
  else if (field >= SALP_I32_BASE && field < SALP_I32_END) { local = field - SALP_I32_BASE; nfields = SALP_NUM_I32; arr = (char*)h->view.i32; *elem = 4; }
  else return SALP_ERR_INVALID;
#if SALP_STATE_AOS
  (void)n;
  *base = arr + *elem * local;
  *stride = *elem * nfields;
#else
  (void)nfields;
  *base = arr + *elem * local * n;
  *stride = *elem;
#endif
  return 0;
}
int salp_get_state(salp_handle h, int32_t field, void* dst, int64_t first, int64_t count) {
  char* base; size_t elem, stride;
  if (!h || column_of(h, field, &base, &elem, &stride)) return SALP_ERR_INVALID;
  for (int64_t k = 0; k < count; k++) memcpy((char*)dst + elem * k, base + stride * (first + k), elem);
  return SALP_OK;
}
int salp_set_state(salp_handle h, int32_t field, const void* src, int64_t first, int64_t count) {
  char* base; size_t elem, stride;
  if (!h || column_of(h, field, &base, &elem, &stride)) return SALP_ERR_INVALID;
  for (int64_t k = 0; k < count; k++) memcpy(base + stride * (first + k), (const char*)src + elem * k, elem);
  return SALP_OK;
}
int salp_trace_cycle(salp_handle h, int64_t env, const float* a, double* trace, int32_t capacity, int32_t* K_out) {
  if (!h || !a || !trace || !K_out || env < 0 || env >= h->view.n) return SALP_ERR_INVALID;
  *K_out = env_trace_cycle(h->params, h->view, env, a[0], a[1], a[2], trace, capacity);
  return SALP_OK;
}
int salp_check(salp_handle h) { return h ? h->status : SALP_ERR_INVALID; }

}  // extern "C"
