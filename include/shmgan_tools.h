/* shmgan_tools.h -- measurement hooks of libshmgan.so.  NOT part of the drop-in boundary (include/shmgan.h): nothing the reference does maps
 * to these; they exist for tools/bench_norm.py and for the device-path equivalence test tests/test_gpu_fullsize.py. */
#ifndef SHMGAN_TOOLS_H_
#define SHMGAN_TOOLS_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Tuning / comparison hook for the instance-norm and activation-backward streams (tools/bench_norm.py): pipe_off = 1 routes bf16 tensors
 * to the register-staged kernels instead of the cp.async-pipelined range kernels; depth in {0 = per-kernel default, 2, 4, 8} is the
 * per-thread ring depth; grid_mul = 0 launches one resident wave, k > 0 launches k blocks per SM.  Defaults: (0, 0, 0). */
int shm_norm_tune(int pipe_off, int depth, int grid_mul);

#ifdef __cplusplus
}
#endif
#endif
