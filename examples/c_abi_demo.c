/* c_abi_demo.c -- the drop-in boundary from plain C: no Python, no torch, only include/wbc_b200.h, libwbc_b200.so and the
 * CUDA runtime.  Reads a case file (tree table, controller configuration, N states with their targets, task memory and
 * references: what the reference holds as a Pinocchio model and as attributes of RobotModel, Robot_Wrapper4.py:19-193),
 * runs ONE batched tick -- everything RobotModel.runWBC does between reading its arguments and returning
 * (Robot_Wrapper4.py:1330-1412) -- through wbc_step and writes qdot / status / iters / q_next to the result file.
 *
 *   gcc -O2 -std=c11 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -L<dir of libwbc_b200.so> -lwbc_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,<dir> -Wl,-rpath,/usr/local/cuda/lib64
 *   ./c_abi_demo case.bin result.bin
 *
 * Case file (little endian): int64 magic 0x57424332, int64 N, double dt, int64 sizeof(WbcTreeTable), the table, int64
 * sizeof(WbcConfig), the config, then q [N, nq], targets [N, 18], mem [N, 72], ref [N, 24] as float64.
 * Result file: qdot [N, nv] f64, q_next [N, nq] f64, status [N] i32, iters [N] i32.
 * tests/test_gpu_surface.py::test_c_abi_demo_matches_python_path writes the case, runs this program and compares bit for bit. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "wbc_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_WBC(x) do { int rc_ = (x); if (rc_ != WBC_OK) { fprintf(stderr, "%s: error %d: %s\n", #x, rc_, wbc_last_error()); return 3; } } while (0)

static int read_exact(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s case.bin result.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int64_t magic = 0, N = 0, sz = 0;
  double dt = 0.0;
  WbcTreeTable* table = (WbcTreeTable*)calloc(1, sizeof(WbcTreeTable));
  WbcConfig* cfg = (WbcConfig*)calloc(1, sizeof(WbcConfig));
  if (read_exact(f, &magic, 8) || magic != 0x57424332 || read_exact(f, &N, 8) || read_exact(f, &dt, 8)) { fprintf(stderr, "bad case header\n"); return 1; }
  if (read_exact(f, &sz, 8) || sz != (int64_t)sizeof(WbcTreeTable) || read_exact(f, table, sizeof(WbcTreeTable))) { fprintf(stderr, "WbcTreeTable size mismatch (%lld vs %zu)\n", (long long)sz, sizeof(WbcTreeTable)); return 1; }
  if (read_exact(f, &sz, 8) || sz != (int64_t)sizeof(WbcConfig) || read_exact(f, cfg, sizeof(WbcConfig))) { fprintf(stderr, "WbcConfig size mismatch (%lld vs %zu)\n", (long long)sz, sizeof(WbcConfig)); return 1; }
  const int nq = table->nq, nv = table->nv;
  const size_t nb_q = (size_t)N * nq * 8, nb_t = (size_t)N * WBC_TARGETS_STRIDE * 8, nb_m = (size_t)N * WBC_MEM_STRIDE * 8;
  const size_t nb_r = (size_t)N * WBC_REF_STRIDE * 8, nb_v = (size_t)N * nv * 8, nb_i = (size_t)N * 4;
  double *hq = malloc(nb_q), *ht = malloc(nb_t), *hm = malloc(nb_m), *hr = malloc(nb_r);
  if (read_exact(f, hq, nb_q) || read_exact(f, ht, nb_t) || read_exact(f, hm, nb_m) || read_exact(f, hr, nb_r)) { fprintf(stderr, "short case file\n"); return 1; }
  fclose(f);

  if (wbc_abi_version() != WBC_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  CHECK_CUDA(cudaSetDevice(0));
  WbcModel* model = NULL;
  CHECK_WBC(wbc_model_create(table, &model));                 /* pin.buildModelFromUrdf + createData (:21-23) */
  double *dq, *dtg, *dm, *dr, *dv, *dqn;
  int32_t *dst, *dit;
  CHECK_CUDA(cudaMalloc((void**)&dq, nb_q)); CHECK_CUDA(cudaMalloc((void**)&dtg, nb_t)); CHECK_CUDA(cudaMalloc((void**)&dm, nb_m));
  CHECK_CUDA(cudaMalloc((void**)&dr, nb_r)); CHECK_CUDA(cudaMalloc((void**)&dv, nb_v)); CHECK_CUDA(cudaMalloc((void**)&dqn, nb_q));
  CHECK_CUDA(cudaMalloc((void**)&dst, nb_i)); CHECK_CUDA(cudaMalloc((void**)&dit, nb_i));
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  CHECK_CUDA(cudaMemcpyAsync(dq, hq, nb_q, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(dtg, ht, nb_t, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(dm, hm, nb_m, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(dr, hr, nb_r, cudaMemcpyHostToDevice, stream));

  WbcStepIO io;
  memset(&io, 0, sizeof(io));
  io.q = dq; io.targets = dtg; io.mem_in = dm; io.ref = dr; io.dt = dt;
  io.qdot = dv; io.status = dst; io.iters = dit;
  io.mem_out = dm;                                            /* task memory advanced in place, as runWBC mutates its object */
  io.q_next = dqn;                                            /* integrate + base estimate (:1397-1402) */
  CHECK_WBC(wbc_step(model, cfg, &io, N, stream));

  double *ov = malloc(nb_v), *oq = malloc(nb_q);
  int32_t *os = malloc(nb_i), *oi = malloc(nb_i);
  CHECK_CUDA(cudaMemcpyAsync(ov, dv, nb_v, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(oq, dqn, nb_q, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(os, dst, nb_i, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(oi, dit, nb_i, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  FILE* g = fopen(argv[2], "wb");
  if (!g) { perror(argv[2]); return 1; }
  fwrite(ov, 1, nb_v, g); fwrite(oq, 1, nb_q, g); fwrite(os, 1, nb_i, g); fwrite(oi, 1, nb_i, g);
  fclose(g);
  long solved = 0; double sum = 0.0;
  for (int64_t s = 0; s < N; ++s) solved += os[s] == 0;
  for (size_t k = 0; k < (size_t)N * nv; ++k) sum += ov[k] < 0 ? -ov[k] : ov[k];
  printf("c_abi_demo: N = %lld, nq = %d, nv = %d: %ld QPs solved, sum |qdot| = %.12g\n", (long long)N, nq, nv, solved, sum);
  wbc_model_destroy(model);
  cudaFree(dq); cudaFree(dtg); cudaFree(dm); cudaFree(dr); cudaFree(dv); cudaFree(dqn); cudaFree(dst); cudaFree(dit);
  cudaStreamDestroy(stream);
  return 0;
}
