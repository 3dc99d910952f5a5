/* shmgan.h -- C ABI of libshmgan.so: the B200 (sm_100a) kernels behind SHMGAN's hot path.
 *
 * The reference (Atif-Anwer/SHMGAN) is TensorFlow/Keras Python with no FFI of its own; the boundary a
 * maintainer would bind is the set of library ops TF dispatches from ShmGANwithSSpecSeg.py / SpecSeg.py
 * (SURVEY.md section 2b).  Each entry point below cites the reference call site it replaces.
 *
 * Conventions
 *   - plain pointers + sizes only; the caller owns every buffer (device memory unless stated);
 *   - activations are NHWC; "ld" = elements between consecutive pixels (>= channels) so that a
 *     tensor may be a channel slice of a wider concat buffer (the slice offset is folded into the pointer);
 *   - dtype: SHM_F32 or SHM_BF16 for activations; weights, gradients of weights, statistics: fp32/fp64;
 *   - weights keep the Keras layouts: Conv2D (kh,kw,Cin,Cout), Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises the host,
 *     allocates nothing, and returns 0 on success or a negative shm_status; the message is in
 *     shm_last_error().  There is NO CPU fallback: unsupported shapes fail loudly.
 */
#ifndef SHMGAN_H_
#define SHMGAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SHM_OK = 0, SHM_EINVAL = -1, SHM_EUNSUPPORTED = -2, SHM_ECUDA = -3 } shm_status;
typedef enum { SHM_F32 = 0, SHM_BF16 = 1 } shm_dtype;
typedef enum { SHM_ACT_NONE = 0, SHM_ACT_LRELU = 1, SHM_ACT_RELU = 2, SHM_ACT_SIGMOID = 3 } shm_act;

/* Geometry of one convolution layer.  H, W are the spatial size of the layer INPUT (x). */
typedef struct {
    int32_t N, H, W;      /* batch, input height, input width                                        */
    int32_t Cin, Cout;
    int32_t kh, kw;       /* 1, 2 or 3                                                               */
    int32_t stride;       /* 1 or 2                                                                  */
    int32_t transposed;   /* 0: Conv2D(padding='same'); 1: Conv2DTranspose(padding='same', strides=2) */
    int32_t act;          /* shm_act fused after the bias in the forward epilogue                    */
    int32_t ldx, ldy;     /* pixel strides (elements) of x / y                                       */
    int32_t dtype;        /* shm_dtype of x, y, dx, dy                                               */
    int32_t tensor_core;  /* 0: exact-fp32 SIMT path (parity mode); 1: tcgen05 bf16 path (needs dtype = SHM_BF16) */
} shm_conv_desc;

const char* shm_last_error(void);
int  shm_version(void);
int  shm_sm_count(void);                 /* SMs of the current device (148 on B200) */
/* zero-fill of a device buffer on `stream` (tf.zeros / tf.zeros_like at ShmGANwithSSpecSeg.py:206,470-471 and the zero-initialised gradient
 * accumulators of tape.gradient :859,:868) */
int  shm_zero(void* ptr, int64_t bytes, void* stream);

/* ---- convolutions: Keras Conv2D ShmGANwithSSpecSeg.py:244,254,263,272,281,301,308,315,322,326,365,387,410-411,
 *      SpecSeg.py:34-88; Conv2DTranspose ShmGANwithSSpecSeg.py:298,305,312,319, SpecSeg.py:64,70,76,82;
 *      dgrad / wgrad = tape.gradient ShmGANwithSSpecSeg.py:859,868 ---- */
int shm_conv2d_fwd  (const shm_conv_desc* d, const void* x, const float* w, const float* bias, void* y, void* stream);
/* dx = dL/dx given dy = dL/d(pre-activation).  accumulate != 0: dx += (fp32 only). */
int shm_conv2d_dgrad(const shm_conv_desc* d, const void* dy, const float* w, void* dx, int accumulate, void* stream);
/* dw += dL/dw (Keras layout, fp32, ALWAYS accumulates: weights are shared by several passes); dbias += column sums (may be NULL) */
int shm_conv2d_wgrad(const shm_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* stream);

/* tensor-core path (tcgen05 implicit GEMM, bf16 in / fp32 accumulate in TMEM / bf16 out).  Weights are re-laid out once per
 * optimiser step as bf16 [tap][n][k] (K-major B operand): for_dgrad = 0 -> (n,k) = (Cout,Cin); 1 -> (Cin,Cout). */
int64_t shm_conv2d_tc_weight_elems(const shm_conv_desc* d);
int shm_conv2d_tc_supported(const shm_conv_desc* d, int for_dgrad);   /* 1 if the shape tiles into 128-point TMA boxes */
/* cin_real (0 = d->Cin): the Keras kernel holds only cin_real < d->Cin input channels; the rest of the bf16 copy is zero (the layer
 * then reads a zero-padded 64-channel input, see shm_pad_channels64) */
int shm_conv2d_tc_prep_weights(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc, int for_dgrad, void* stream);
/* Forward + instance-norm statistics in ONE kernel ("IN stats fused into the producing conv's epilogue"; the Conv -> LeakyReLU ->
 * InstanceNormalization blocks of ShmGANwithSSpecSeg.py:244-245, :386-389): stats[n][c] = (sum, sum of squares) over the pixels of image n of
 * the bf16 values stored to y, ADDED to the caller's zero-initialised fp64 buffer [N][Cout][2] -- exactly what shm_inorm_stats(y) returns,
 * without reading y back.  shm_conv2d_tc_stats_supported: 1 if the kernel serving this layer has the statistics epilogue. */
int shm_conv2d_tc_stats_supported(const shm_conv_desc* d);
int shm_conv2d_tc_fwd_stats(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, double* stats, void* stream);
/* which tcgen05 kernel serves the layer (accounting only): pass 0 fwd, 1 dgrad, 2 wgrad -> 0 conv_tc, 1 conv_halo, 2 conv_multi (big),
 * 3 conv_multi (stride-2 scatter), 4 wgrad_tc, 5 wgrad_halo<0>, 6 wgrad_halo<1>, 7 wgrad_s2<0>, 8 wgrad_s2<1>; -1 = not servable */
int shm_conv2d_tc_route(const shm_conv_desc* d, int pass);
/* both layouts (fwd and dgrad) in one launch */
int shm_conv2d_tc_prep_weights_both(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc_fwd, void* w_tc_dgrad, void* stream);
/* Batched form of shm_conv2d_tc_prep_weights_both for a whole network: fill one job record per layer in HOST memory
 * (shm_conv2d_tc_prep_job_bytes() bytes each, contiguous), finalize the array (returns the launch's block count), copy it to the
 * device once; shm_conv2d_tc_prep_multi then refreshes every layer's two bf16 layouts in ONE launch per optimiser step. */
int shm_conv2d_tc_prep_job_bytes(void);
int shm_conv2d_tc_prep_job(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc_fwd, void* w_tc_dgrad, void* job_out);
int shm_conv2d_tc_prep_jobs_finalize(void* jobs_host, int njobs);
int shm_conv2d_tc_prep_multi(const void* jobs_dev, int njobs, int total_blocks, void* stream);
/* forward-layout weights of a layer run in a zero-padded device geometry (d->Cin, d->Cout) >= the Keras kernel's (cin_real, cout_real):
 * device input channel k holds real channel (k / seg_pad) * seg_real + k %% seg_pad when k %% seg_pad < seg_real, zero otherwise
 * (one segment = zero-padded input; two = concat of two zero-padded halves, SpecSeg.py:65-83 at 16/32 channels) */
int shm_conv2d_tc_prep_weights_padded(const shm_conv_desc* d, const float* w, int cin_real, int seg_real, int seg_pad, int cout_real,
                                      void* w_tc, void* stream);
int shm_conv2d_tc_fwd  (const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, void* stream);
/* Same, storing only the first `nstore` (multiple of 8) output channels of every pixel: the remaining Cout - nstore columns are the zero
 * padding of a layer whose real channel count is below the tensor-core granule (SpecSeg.py:68,78: Conv2DTranspose to 32 / 16 channels
 * written into the [up | skip] concat buffer).  Served by the Conv2DTranspose scatter kernel only (SHM_EUNSUPPORTED otherwise). */
int shm_conv2d_tc_fwd_cols(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, int nstore, void* stream);
int shm_conv2d_tc_dgrad(const shm_conv_desc* d, const void* dy, const void* w_tc_dgrad, void* dx, void* stream);
int shm_conv2d_tc_wgrad(const shm_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);   /* dw += (fp32 atomics) */
/* dbias[c] += sum over pixels of dy[pix, c] */
int shm_colsum(const void* dy, int64_t npix, int C, int ld, int dtype, float* out, void* stream);

/* ---- instance norm: tfa InstanceNormalization ShmGANwithSSpecSeg.py:245,...,388 (Generator_summary.txt:9-37) ---- */
/* sums[N][C][2] (fp64) += (sum x, sum x^2) over H*W; caller zeroes sums first */
int shm_inorm_stats(const void* x, int N, int HW, int C, int ldx, int dtype, double* sums, void* stream);
/* y = (x-mean)*rstd*gamma+beta (+ add).  out (ld ldo) and/or pooled = AvgPool2x2(y without add) (ld ldp) may be NULL.
 * AveragePooling2D ShmGANwithSSpecSeg.py:249; skip blend :290-293; D blend :359 */
int shm_inorm_apply(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums,
                    const float* gamma, const float* beta, float eps,
                    const void* add, int ldadd, int nadd /* add has nadd images, broadcast as n %% nadd; 0 = N */,
                    void* out, int ldo, void* pooled, int ldp, void* stream);
/* backward: dy = dyA (ld ldA, may be NULL) + 0.25 * upsample2(dyP) (may be NULL).
 * pass 1 accumulates bsums[N][C][2] (fp64) += (sum dy, sum dy*xhat); pass 2 writes
 * dx = act'(x) * rstd*gamma*(dy - mean(dy) - xhat*mean(dy*xhat)), x being the saved POST-activation conv output. */
int shm_inorm_bwd_stats(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums, float eps,
                        const void* dyA, int ldA, const void* dyP, int ldP, double* bsums, void* stream);
int shm_inorm_bwd_apply(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums,
                        const float* gamma, float eps, const void* dyA, int ldA, const void* dyP, int ldP,
                        const double* bsums, int act, void* dx, int lddx,
                        float* dbias /* may be NULL: dbias[c] += sum over pixels of dx = the producing conv's bias gradient */, void* stream);

/* ---- pointwise ---- */
/* dpre = dy * act'(y_post)   (LeakyReLU/ReLU derivative from the saved post-activation value); dbias (may be NULL) += column sums of dpre */
int shm_act_bwd(const void* dy, int lddy, const void* y, int ldy, void* dpre, int ldd, int64_t npix, int C, int act, int dtype, float* dbias, void* stream);
/* MaxPooling2D(k) ShmGANwithSSpecSeg.py:406 (k=2), :358 (k=16); SpecSeg.py:38 */
int shm_maxpool(const void* x, int N, int H, int W, int C, int ldx, int k, void* y, int ldy, int dtype, void* stream);
/* Keras BatchNormalization at predict time (SpecSeg.py:37): y = (x-mean)*rsqrt(var+eps)*gamma+beta; optional fused MaxPool2 */
int shm_bn_eval(const void* x, int N, int H, int W, int C, int ldx, int dtype, const float* gamma, const float* beta,
                const float* mean, const float* var, float eps, void* out, int ldo, void* pooled, int ldp, void* stream);
/* out = a + b (strided; b may have ld) ; used for GaussianNoise (:352) */
int shm_add(const void* a, int lda, const void* b, int ldb, void* out, int ldo, int64_t npix, int C, int dtype, void* stream);
/* Dropout (:363): out = x * keep * scale */
/* dst[b,pix,c] (+)= sum_r src[r*nb + b, pix, c]: gradient of a batch-broadcast add (mask attention shared by all passes) */
int shm_group_sum(const void* src, int lds, int reps, int64_t pix_per_group, int C, void* dst, int ldd, int accumulate, int dtype, void* stream);
int shm_mul_mask(const void* x, const void* keep, void* out, int64_t n, float scale, int dtype, void* stream);
/* counter-based RNG (Philox-4x32-10): normal(0, sigma) / Bernoulli keep mask */
int shm_rng_normal(void* out, int64_t n, uint64_t seed, uint64_t offset, float sigma, int dtype, void* stream);
int shm_rng_keep(void* out, int64_t n, uint64_t seed, uint64_t offset, float keep_prob, int dtype, void* stream);
/* the same with the counter offset read from device memory when the kernel runs: a train_step captured in a CUDA graph (model.py, cuda_graph=True)
 * replays with a fresh GaussianNoise / Dropout stream every step (:352, :363) */
int shm_rng_normal_dev(void* out, int64_t n, uint64_t seed, const uint64_t* offset_dev, float sigma, int dtype, void* stream);
int shm_rng_keep_dev(void* out, int64_t n, uint64_t seed, const uint64_t* offset_dev, float keep_prob, int dtype, void* stream);
int shm_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* strided copy / convert of an [npix, C] plane set: dst[p*ldd + c] = src[p*lds + c]  (Y-channel extraction :486-490, batch stacking) */
int shm_cast2d(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, int64_t npix, int C, void* stream);
int shm_axpy(float alpha, const void* x, void* y, int64_t n, int dtype, void* stream);   /* y += alpha x */

/* ---- Dense (ShmGANwithSSpecSeg.py:371-375): x [B,K] (Flatten of NHWC), w [K,J] fp32 ---- */
int shm_dense_fwd  (const void* x, const float* w, float* out, int B, int K, int J, int dtype, void* stream);  /* out zeroed by callee */
int shm_dense_dgrad(const float* dout, const float* w, void* dx, int B, int K, int J, int dtype, void* stream);
int shm_dense_wgrad(const void* x, const float* dout, float* dw, int B, int K, int J, int dtype, void* stream); /* dw += */

/* 1x1 convolution to ONE output channel (generator output ShmGANwithSSpecSeg.py:326, SpecSeg head SpecSeg.py:88) as a bandwidth kernel.
 * bf16 activations, C in {8,16,32,64,128,256}; w [C] fp32 (the Keras (1,1,C,1) kernel), bias [1] or NULL.
 * bwd fuses the activation derivative (from the saved post-activation y): dx (may be NULL) = dpre * w, dw += x^T dpre, dbias += sum dpre */
int shm_pw1_fwd(const void* x, int ldx, int C, const float* w, const float* bias, int act, void* y, int64_t npix, int dtype, void* stream);
int shm_pw1_bwd(const void* x, int ldx, int C, const float* w, const void* dy, const void* y, int act, void* dx, int lddx,
                float* dw, float* dbias, int64_t npix, int dtype, void* stream);

/* 3x3 stride-1 SAME convolution to ONE output channel (the discriminator's real/fake head ShmGANwithSSpecSeg.py:365-369) as
 * warp-per-pixel dot products.  bf16 activations, C %% 8 == 0; w = the Keras (3,3,C,1) kernel (fp32 [9][C]); y / dpre are [N,H,W] bf16.
 * dgrad / wgrad take dpre = dL/d(pre-activation) (see shm_act_bwd); dw += */
int shm_c3to1_fwd(const void* x, int N, int H, int W, int C, int ldx, const float* w, const float* bias, int act, void* y, int dtype, void* stream);
int shm_c3to1_dgrad(const void* dpre, int N, int H, int W, int C, const float* w, void* dx, int lddx, int dtype, void* stream);
int shm_c3to1_wgrad(const void* x, int N, int H, int W, int C, int ldx, const void* dpre, float* dw, int dtype, void* stream);

/* ---- polarimetric preprocessing ---- */
/* calculate_estimate_diffuse utils.py:102-106: per-element min of four images (n elements). dtype: 0 f32, 1 bf16, 2 u8 */
int shm_pseudo_diffuse_min4(const void* i0, const void* i45, const void* i90, const void* i135, void* out, int64_t n, int dtype, void* stream);
/* rgb_to_yuv + custom_per_image_standardization ShmGANwithSSpecSeg.py:480-484, :1271-1309.
 * pass 1: sums[N][2] (fp64) += (sum v, sum v^2) over the image's yuv values; pass 2: yuv = rgb2yuv(rgb)/scale;
 * scale[N] (fp32) = max(std, 1/256) is also written. rgb is fp32 (dataset tensors), yuv fp32. */
int shm_yuv_stats(const float* rgb, int N, int HW, double* sums, void* stream);
int shm_yuv_standardize(const float* rgb, int N, int HW, const double* sums, float* yuv, float* scale, void* stream);
/* averageCbCr ShmGANwithSSpecSeg.py:505 */
int shm_avg_cbcr(const float* y0, const float* y1, const float* y2, const float* y3, const float* y4, float* out, int64_t npix, void* stream);
/* generator input assembly :509-531 / :576-594 / test.py:227-235.  For slot j: src[j] (fp32, pixel stride src_ld[j]) or NULL = zeros.
 * out [npix,10] (dtype): 5 slots then the one-hot plane `onehot`. */
int shm_assemble_input(const float* const src[5], const int32_t src_ld[5], int onehot, void* out, int ldo /* >= 10; channels 10..ldo-1 are zero-filled */, int64_t npix, int dtype, void* stream);
/* dst (bf16, pixel stride 64) = src channels 0..C-1 (C <= 64) followed by zeros: the zero-padded input that lets the first layers
 * (Cin = 10 / 3 / 1) run on the tensor-core kernels, whose reduction dimension moves in 64-channel TMA boxes */
int shm_pad_channels64(const void* src, int src_dtype, int lds, int C, void* dst_bf16, int64_t npix, void* stream);
/* same with a padded width Cpad in {16, 32, 64}: the inputs of the thin tensor-core layers (16 / 32 reduction channels per pixel row) */
int shm_pad_channels(const void* src, int src_dtype, int lds, int C, void* dst_bf16, int Cpad, int64_t npix, void* stream);
/* im2col of a 3x3 stride-2 SAME conv on a few-channel image (discriminator d1, ShmGANwithSSpecSeg.py:353): out bf16 [N,H/2,W/2,64],
 * channel (ky*3+kx)*C + c = x[n, 2oy+ky-pb, 2ox+kx-pb, c] (zero outside / beyond 9*C); d1 then runs as a 1x1 conv on the tensor cores.
 * col2im is its transpose (the d(image) of the generator-loss path). */
int shm_im2col_k3s2(const void* x, int src_dtype, int ldx, int N, int H, int W, int C, void* out_bf16, void* stream);
int shm_col2im_k3s2(const void* dP_bf16, int N, int H, int W, int C, void* dx, int dst_dtype, int lddx, void* stream);
/* backward of the cyclic assembly: dgen[npix] += sum over listed slots of din[npix,ldin][slot] */
int shm_assemble_bwd(const void* din, int dtype, int ldin, const int32_t slots[5], int nslots, float* dgen, int64_t npix, void* stream);
/* yuv_to_rgb(concat(Y, CbCr)) :544,553,613-624.  Y fp32 (ld 1), cbcr fp32 [npix,2]; rgb out fp32 and / or a copy in dtype_lp with pixel
 * stride ld_lp (channels 3..ld_lp-1 zero-filled) for the discriminator */
int shm_yuv2rgb(const float* Y, const float* cbcr, int64_t npix_cbcr /* cbcr index = pixel %% npix_cbcr */, float* rgb, void* rgb_lp, int dtype_lp, int ld_lp, int64_t npix, void* stream);
/* dY (+)= sum_c (drgb_f32 + drgb_lp)[.,c]  (d rgb / dY = (1,1,1)); either gradient source may be NULL */
int shm_yuv2rgb_bwd(const float* drgb_f32, const void* drgb_lp, int dtype_lp, int ld_lp, float* dY, int64_t npix, int accumulate, void* stream);

/* running mean of the standardisation scales: acc[0] += sum(src[0..n)), acc[1] += n (fp64, device).  Replaces the unbounded
 * self.stddev_arr list (ShmGANwithSSpecSeg.py:1306, datasetLoader.py:42) whose tf.reduce_mean is read at :548 and test.py:246 */
int shm_sum_count(const float* src, int64_t n, double* acc, void* stream);
/* out = x * mul * (acc ? acc[0]/acc[1] : 1): gen_rgb_output = yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255) (:550, test.py:249; yuv_to_rgb is linear) */
int shm_scale_by_mean(const float* x, const double* acc, float mul, float* out, int64_t n, void* stream);

/* ---- losses (ShmGANwithSSpecSeg.py:669-844).  All accumulate `weight * loss` into loss_out[0] (fp32, device) and write gradients ---- */
/* mean((a - target)^2) over n; da (+)= gscale * 2 (a-target)/n.  :669-679, :721-728 */
int shm_lsgan(const float* a, int64_t n, float target, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream);
/* mean_b softmax-CE(labels[5], logits[b]) :695-714; dlogits (+)= gscale*(sum(labels)*softmax - labels)/B */
int shm_softmax_ce(const float* logits, int B, const float labels[5], float* loss_out, float weight, float* dlogits, float gscale, int accumulate, void* stream);
/* the same two with the target / the 5 labels read from device memory when the kernel runs: TARGET_LABELS is redrawn every step (:986), and a
 * train_step captured in a CUDA graph (model.py, cuda_graph=True) must see the current draw */
int shm_lsgan_dev(const float* a, int64_t n, const float* target_dev, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream);
int shm_softmax_ce_dev(const float* logits, int B, const float* labels5_dev, float* loss_out, float weight, float* dlogits, float gscale, int accumulate, void* stream);
/* mean|a-b| :744-751;  da (+)= gscale*sign(a-b)/n */
int shm_l1(const float* a, const float* b, int64_t n, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream);
/* mean (a-b)^2 */
int shm_mse(const float* a, const float* b, int64_t n, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream);
/* content loss :814 on concat(Y, cbcr) vs yuv [npix,3]; dY += gscale * 2 (Y - yuv0) / (3 npix) */
int shm_mse_ycc(const float* Y, const float* cbcr, const float* yuv, int64_t npix, float* loss_out, float weight, float* dY, float gscale, void* stream);
/* per-image min/max over [HW*3] of concat(Y, cbcr) or of a plain [HW,3] tensor (cbcr == NULL): mm[N][2] = {min,max}; idx[N][2] = flat argmin/argmax (pix*3+c) */
int shm_minmax3(const float* Y_or_yuv, const float* cbcr, int N, int HW, float* mm, int32_t* idx, void* stream);
/* Gram 3x3 per image (:1176-1180) of concat(Y,cbcr) or plain yuv: gram[N][9] (fp64) += sum_p x_c x_d (the 1/HW is applied by the consumers); caller zeroes */
int shm_gram3(const float* Y_or_yuv, const float* cbcr, int N, int HW, double* gram, void* stream);
/* style loss :817-821 from two gram buffers; dgram[N][9] (fp32) = d(weight*style)/dgramA */
int shm_style_loss(const double* gramA, const double* gramB, int N, int HW, int S, float* loss_out, float weight, float* dgramA, float gscale, void* stream);
/* dY[n,p] += sum_d (dgram[c=0,d] + dgram[d,0]) * x_d / HW  (only the Y channel carries gradient to G) */
int shm_gram3_bwd(const float* Y, const float* cbcr, int N, int HW, const float* dgram, float* dY, void* stream);
/* tf.image.ssim on rescale_01'd images (:759-779, utils.py:190-195).  imgA = concat(Y,cbcr) (generated), imgB = yuv [N,HW,3] (reference image).
 * ssim_out[N]; maps[N][3][Ho][Wo][3] saved partials for the backward (may be NULL to skip).
 * cbcr == NULL: Y is a packed [N,HW,3] image (the value-only form the test-time metric test.py:335 uses; maps must be NULL). */
int shm_ssim_fwd(const float* Y, const float* cbcr, const float* mmA, const float* imgB, const float* mmB, int N, int H, int W,
                 float max_val, float* ssim_out, float* maps, void* stream);
/* loss term: weight * mean_b(-log((1+ssim_b)/2)); dssim[N] = gscale * d/dssim_b */
int shm_ssim_loss(const float* ssim, int N, float* loss_out, float weight, float* dssim, float gscale, void* stream);
/* dY += d(loss)/dY through ssim and rescale_01 (including the min / max paths) */
int shm_ssim_bwd(const float* Y, const float* cbcr, const float* mmA, const int32_t* idxA, const float* imgB, const float* mmB, int N, int H, int W,
                 const float* maps, const float* dssim, float* dY, double* scratch /* [N][2] */, void* stream);
int64_t shm_ssim_map_elems(int N, int H, int W);   /* floats needed for `maps` */
/* masked L2 "Spec" term :792-806 (value only; it is not part of any total) */
int shm_spec_loss(const float* Y, const float* cbcr, const float* yuv, const float* mask, int64_t npix, float* loss_out, float weight, void* stream);

/* ---- clip_by_value(+-clip) + Keras Adam (:860-871, optimizers :169-175) over a flat parameter buffer.
 *      lr_t = lr(step)*sqrt(1-b2^t)/(1-b1^t) is computed by the caller.  gscale multiplies the gradient BEFORE the clip
 *      (1/world for the data-parallel average).  */
int shm_clip_adam(float* param, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
                  float eps, float clip, float gscale, void* stream);
/* lr_t read from device memory when the kernel runs (CUDA-graph replays of train_step: the ExponentialDecay / bias-correction factor of :169-175 moves every step) */
int shm_clip_adam_dev(float* param, const float* grad, float* m, float* v, int64_t n, const float* lr_t_dev, float beta1, float beta2,
                      float eps, float clip, float gscale, void* stream);

/* ---- rows SURVEY.md 8(f) marks "next": loader contract, degree of polarisation, test-time metrics (csrc/extras.cu) ---- */
/* datasetLoader.py:48-62: decoded uint8 images src [N,Hs,Ws,3] -> tf.image.resize(bilinear, half-pixel centres) to [N,Ho,Wo,3] fp32,
 * x / 255.0 (:60), rows reversed when flip_ud != 0 (tf.image.flip_up_down, :61 -- the reference flips when random_flip is False).
 * Bit-exact with the float32 order of operations of TF's ResizeBilinear kernel. */
int shm_load_u8_bilinear(const void* src, int N, int Hs, int Ws, float* dst, int Ho, int Wo, int flip_ud, void* stream);
/* calcDOP ShmGANwithSSpecSeg.py:1157-1169: dop = divide_no_nan(sqrt((I0-I90)^2 + (I45-I135)^2), I0+I90); aop (may be NULL) = 0.5 atan2(S2, S1) */
int shm_dop(const float* i0, const float* i45, const float* i90, const float* i135, float* dop, float* aop, int64_t n, void* stream);
/* test.py:338,343 (tf.image.psnr / MeanSquaredError): out[n] += sum over the `per` elements of image n of (a - b)^2, fp64 */
int shm_sqerr_per_image(const float* a, const float* b, int N, int64_t per, double* out, void* stream);
/* test.py:346-349: sRGB -> Lab (tfio rgb_to_lab: D65, 2 degree) of both images, then sums[n][0] += sum dE76, sums[n][1] += sum dE94
 * (skimage deltaE_cie76 / deltaE_ciede94 defaults, rgb1 = the reference colour).  rgb [N,HW,3] fp32 in [0,1]. */
int shm_delta_e(const float* rgb1, const float* rgb2, int N, int64_t HW, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHMGAN_H_ */
