/* abl_b200.h -- C ABI of libabl_b200.so: the B200-native (sm_100a) numerics of the cutoffvgg16
 * hot path of gchetty/ab-line-classifier.
 *
 * The reference has NO native code: every numeric op on this path is a TensorFlow 2.9 / Keras
 * library call made from Python.  Each entry point below therefore cites the
 * reference call site whose Keras op(s) it replaces; the binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md), mirrored by ab_line_classifier_b200/_native.py.
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller; nothing here allocates or frees.
 *   - Activations are NHWC, bf16 unless stated; master weights/gradients are fp32 in Keras layout
 *     (conv kernels HWIO [3][3][Cin][Cout], dense (in,out)).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and stream-ordered.
 *   - Return value: 0 on success, negative ABL_ERR_* otherwise; abl_last_error() (thread-local)
 *     describes the failure.  No exceptions cross this boundary.
 *   - Re-entrant; no global mutable state besides per-device one-time kernel attributes.
 *   - Deterministic: no floating-point atomics anywhere.  Every sum across CTAs goes through per-CTA partial sums
 *     that a later kernel adds in a fixed order, so two runs on the same inputs give the same bits (gradients, loss,
 *     weights after any number of steps).
 */
#ifndef ABL_B200_H_
#define ABL_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ABL_ABI_VERSION 8

#define ABL_OK 0
#define ABL_ERR_SHAPE (-1)
#define ABL_ERR_ALIGN (-2)
#define ABL_ERR_NULL (-3)
#define ABL_ERR_DRIVER (-4)
#define ABL_ERR_CUDA (-5)

/* epilogue flags of abl_conv3x3_igemm */
#define ABL_EPI_BIAS_RELU 1 /* y = relu(acc + bias[co])                  (forward)              */
#define ABL_EPI_RELU_MASK 2 /* y = mask_src[n,h,w,co] > 0 ? acc : 0      (dgrad -> dZ of layer below) */

int abl_abi_version(void);
const char* abl_last_error(void);

/* Partial-sum ("parts") buffers -- how a bias gradient travels from the kernel that produces a dZ tensor to the kernel
 * that finishes it, without atomics and without re-reading dZ.  A parts buffer is caller-owned fp32 device memory,
 * 16-byte aligned, of at least abl_parts_bytes(C) bytes:
 *   word 0                           int32 R = number of rows the producer wrote (one per CTA)
 *   floats [ABL_PARTS_HEADER + r*C .. +C)   row r: that CTA's partial sums of the C channels
 * Producers (the *_parts arguments below) overwrite the buffer; consumers (abl_conv3x3_wgrad's db_parts,
 * abl_parts_reduce) add rows 0..R-1 in a fixed order. */
#define ABL_PARTS_HEADER 64
long long abl_parts_bytes(int C);
/* out[c] (+)= sum_r row_r[c]  (accumulate != 0 adds to out, else overwrites it); C % 4 == 0. */
int abl_parts_reduce(const float* parts, float* out, int C, int accumulate, void* stream);

/* Conv2D 3x3/s1/SAME as a tcgen05 implicit GEMM.  Replaces Keras Conv2D+BiasAdd+Relu(+MaxPool2D)
 * of the VGG16 layers reused at src/models/models.py:409-413 (forward) and, with the tap-flipped
 * transposed filter and ABL_EPI_RELU_MASK, TF's Conv2DBackpropInput+ReluGrad that model.fit's
 * GradientTape runs (models.py:427,432).
 *   x        bf16 [B,H,W,Cin]
 *   w_kmajor bf16 [9][Cout][Cin]   (from abl_prep_conv_weights: w_fwd for forward, w_dgrad for dgrad;
 *                                   taps kw-major)
 *   bias     fp32 [Cout]           (ABL_EPI_BIAS_RELU)
 *   mask_src bf16 [B,H,W,Cout]     (ABL_EPI_RELU_MASK; post-ReLU activation of the layer below)
 *   out_full bf16 [B,H,W,Cout] or NULL;  out_pool bf16 [B,H/2,W/2,Cout] (2x2/s2 max) or NULL
 *   pool_code uint16 [B,H/2,W/2,Cout/8] or NULL (needs out_pool): per pooled element the 2-bit index (window scan
 *            order) of the FIRST maximum, channel k of an 8-channel chunk in bits [2k,2k+1] -- what MaxPoolGrad needs,
 *            so training keeps neither the full-resolution activation nor re-reads it (abl_maxpool_bwd_code)
 *   colsum_parts  parts buffer (see above) or NULL: partial sums over (n,h,w) of out_full[n,h,w,c] -- with the dgrad
 *            epilogue this is TF's BiasAddGrad of the layer below, fused so its dZ is not read a second time (needs
 *            out_full); colsum_parts_bytes = its capacity
 * Requires Cin, Cout multiples of 64 (<= 512).  out_pool is [B, H/2, W/2, Cout] with the divisions rounding DOWN (Keras
 * MaxPooling2D, padding 'valid': an odd side loses its last row / column).  ANY H x W: the kernels tile a frame into 8 x 16
 * (or 16 x 16) pixel rectangles counted rounding up -- TMA zero-fills loads beyond the frame (= SAME padding) and clips
 * the stores, and the epilogue keeps rows outside the frame out of the pooled output, the routing codes and every sum. */
int abl_conv3x3_igemm(const void* x, const void* w_kmajor, const float* bias, const void* mask_src,
                      void* out_full, void* out_pool, unsigned short* pool_code, float* colsum_parts,
                      long long colsum_parts_bytes, int B, int H, int W, int Cin, int Cout, int flags, void* stream);

/* Forward conv + bias + ReLU (as abl_conv3x3_igemm with ABL_EPI_BIAS_RELU) of the LAST conv layer with Keras'
 * GlobalAveragePooling2D (models.py:414) fused in: the epilogue also writes the per-channel sums of every 8x16-pixel
 * rectangle it produces, gap_part fp32 [B * parts][Cout] (one writer per value: deterministic), and abl_gap_reduce adds
 * a frame's rectangles in fixed order.  parts = rectangles per frame, tiles counted rounding up:
 *   Cout % 256 == 0: ceil(H/8) * ceil(W/16);   otherwise (16-row tiles): 2 * ceil(H/16) * ceil(W/16)
 * (pixels outside the frame add 0).  out_full may be NULL (inference without Grad-CAM: the activation never reaches
 * HBM).  relu_bits (or NULL): uint32 [B,H,W,Cout/32], bit k of word g = (output channel 32 g + k is > 0) -- the 1-bit
 * ReLU mask of the activation.  A training step passes relu_bits and out_full = NULL: after the GAP sums nothing but
 * the sign of this layer's activation is used (ReluGrad in abl_gap_bwd_bits), so 16 bits per element are neither written
 * nor read back.  Requires Cout % 128 == 0 (relu_bits: Cout % 256 == 0); any H, W. */
int abl_conv3x3_igemm_gap(const void* x, const void* w_kmajor, const float* bias, void* out_full, float* gap_part,
                          unsigned int* relu_bits, int B, int H, int W, int Cin, int Cout, void* stream);
/* gap[n][c] = (1/HW) * sum_{p < parts} gap_part[n*parts + p][c]   (parts as defined above) */
int abl_gap_reduce(const float* gap_part, float* gap, int B, int parts, int C, int HW, void* stream);

/* Conv2DBackpropInput + ReluGrad + MaxPoolGrad (+ BiasAddGrad) in ONE kernel, for a conv whose input is the output of a
 * 2x2 max-pool: the dgrad GEMM at the pooled resolution H x W, relu'(pooled) applied in the epilogue, and every value
 * routed straight to the window element the forward's routing code names (zeros elsewhere), so the pooled-resolution
 * gradient never reaches HBM and no separate MaxPoolGrad pass runs.
 *   dz bf16 [B,H,W,Cin_gemm] (gradient of this conv's output); w_dgrad bf16 [9][Cout][Cin_gemm] (abl_prep_conv_weights);
 *   pooled bf16 [B,H,W,Cout], pool_code uint16 [B,H,W,Cout/8] (both from the forward abl_conv3x3_igemm of the layer
 *   below); dz_out bf16 [B,2H,2W,Cout]; db_parts: parts buffer or NULL receiving the partial sums of the routed
 *   gradients (BiasAddGrad of the layer below).  Requires Cout <= 128 (one N tile); any H, W. */
int abl_conv3x3_dgrad_unpool(const void* dz, const void* w_dgrad, const void* pooled, const unsigned short* pool_code,
                             void* dz_out, float* db_parts, long long db_parts_bytes, int B, int H, int W, int Cin,
                             int Cout, void* stream);

/* Conv2DBackpropFilter (+BiasAddGrad) of model.fit's tape (models.py:427,432) as a tcgen05 GEMM with
 * MN-major operands read straight from the NHWC tensors, split over pixel slices into `workspace`
 * and reduced in a fixed order (deterministic), fp32 accumulation:
 *   dw[kh][kw][ci][co] (+)= sum_{n,h,w} x[n,h+kh-1,w+kw-1,ci] * dz[n,h,w,co]   (HWIO fp32)
 *   db[co]               = sum_{n,h,w} dz[n,h,w,co]                            (if db != NULL; overwrites)
 * x bf16 [B,H,W,Cin], dz bf16 [B,H,W,Cout]; accumulate != 0 adds to dw, else overwrites it.  Cin = 64 or a multiple
 * of 128, Cout a multiple of 64, both <= 512; any H, W (4 x 16-pixel tiles counted rounding up, zero-filled by TMA).
 * db_parts: the parts buffer that the kernel which PRODUCED dz filled (abl_conv3x3_igemm colsum_parts,
 *   abl_conv3x3_dgrad_unpool, abl_gap_bwd_relu, abl_maxpool_bwd_*): the slice-reduction kernel of this call adds its rows
 *   into db.  NULL with db != NULL: the column sums come from this call itself (Cin == 64: a ones row of the GEMM; else a
 *   pass over dz).
 * workspace: >= abl_conv3x3_wgrad_workspace_bytes(...) bytes, 16-byte aligned, caller-owned. */
long long abl_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int abl_conv3x3_wgrad(const void* x, const void* dz, float* dw, float* db, const float* db_parts, void* workspace,
                      long long workspace_bytes, int accumulate, int B, int H, int W, int Cin, int Cout,
                      void* stream);
/* BiasAddGrad alone, first half: db_parts rows = per-block sums over rows of dz bf16 [rows][C] (finish with
 * abl_parts_reduce or abl_conv3x3_wgrad's db_parts). */
int abl_bias_grad(const void* dz, float* db_parts, long long db_parts_bytes, long long rows, int C, void* stream);

/* block1_conv1 (Cin=3, K=27: HBM-bound, not tensor-core shaped).  Forward reads the caller's fp32
 * NHWC frame (already BGR / mean-subtracted, preprocessor.py:60-63) and writes bf16 relu(conv+b).  Any B, H, W (tiles
 * are 128 raster pixels; the last one may be short). */
int abl_conv_first_fwd(const float* x, const float* w_hwio, const float* bias, void* out, int B, int H, int W,
                       int Cout, void* stream);
/* dw[3][3][3][Cout] = ..., db[Cout] = ... (overwritten; db may be NULL) for block1_conv1 (x fp32 [B,H,W,3], dz bf16
 * [B,H,W,Cout]): per-CTA sums go to `workspace` (abl_conv_first_wgrad_workspace_bytes() bytes) and a second kernel adds
 * them in a fixed order. */
long long abl_conv_first_wgrad_workspace_bytes(void);
int abl_conv_first_wgrad(const float* x, const void* dz, float* dw, float* db, void* workspace,
                         long long workspace_bytes, int B, int H, int W, int Cout, void* stream);

/* Same two kernels fed by the RAW uint8 RGB frame [B,H,W,3] (what decode_jpeg + nearest resize leave,
 * reference preprocessor.py:76-80): Keras vgg16 preprocess_input (RGB->BGR, subtract 103.939/116.779/123.68;
 * models.py:13, preprocessor.py:60-63) is fused into the gather, so the fp32 tensor never exists and the
 * host->device copy is 4x smaller. */
int abl_conv_first_fwd_u8(const unsigned char* frames_rgb, const float* w_hwio, const float* bias, void* out, int B,
                          int H, int W, int Cout, void* stream);
int abl_conv_first_wgrad_u8(const unsigned char* frames_rgb, const void* dz, float* dw, float* db, void* workspace,
                            long long workspace_bytes, int B, int H, int W, int Cout, void* stream);

/* fp32 HWIO master kernel -> the two bf16 GEMM operand layouts, taps indexed kw-major (t' = kw*3 + kh):
 *   w_fwd  [kw*3+kh][co][ci]         = w[kh][kw][ci][co]
 *   w_dgrad[(2-kw)*3+(2-kh)][ci][co] = w[kh][kw][ci][co]   (180-degree rotated)      (either may be NULL) */
int abl_prep_conv_weights(const float* w_hwio, void* w_fwd, void* w_dgrad, int Cin, int Cout, void* stream);
/* The same for n_layers (<= 16) layers in ONE launch (after each optimizer step); the five arrays are HOST
 * arrays of n_layers entries (device pointers / channel counts).  Cin and Cout must be multiples of 32 (32x32 tiles
 * are transposed through shared memory so that the source and both copies are accessed coalesced). */
int abl_prep_conv_weights_multi(const float* const* w_hwio, void* const* w_fwd, void* const* w_dgrad,
                                const int* cin, const int* cout, int n_layers, void* stream);

/* MaxPoolGrad + ReluGrad: dz[n,h,w,c] = dp[n,h/2,w/2,c] if a[n,h,w,c] is the first maximum of its 2x2
 * window and a > 0, else 0.  a bf16 [B,H,W,C] (pre-pool, post-ReLU), pooled = maxpool(a), dp, dz bf16.  Odd H or W:
 * dp / pooled are [B,H/2,W/2,C] rounding down, and the last row / column of dz (no window) is zero.
 * db_parts: parts buffer or NULL receiving the partial sums of dz[..., c] (BiasAddGrad of the conv above the pool). */
int abl_maxpool_bwd_relu(const void* dp, const void* a, const void* pooled, void* dz, float* db_parts,
                         long long db_parts_bytes, int B, int H, int W, int C, void* stream);

/* MaxPoolGrad from the routing codes of abl_conv3x3_igemm (pool_code): dz[n,h,w,c] = dp[n,h/2,w/2,c] for the window
 * element the code names, else 0.  dp must already be masked by relu'(pooled) (run the dgrad that produces it with
 * ABL_EPI_RELU_MASK and mask_src = the pooled activation).  db_parts as in abl_maxpool_bwd_relu. */
int abl_maxpool_bwd_code(const void* dp, const unsigned short* pool_code, void* dz, float* db_parts,
                         long long db_parts_bytes, int B, int H, int W, int C, void* stream);

/* GlobalAveragePooling2D (models.py:414): gap[n][c] = mean_{hw} a[n,hw,c]; a bf16, gap fp32. */
int abl_gap_fwd(const void* a, float* gap, int B, int HW, int C, void* stream);
/* its gradient fused with ReluGrad of the last conv: dz = a > 0 ? dgap[n][c] / HW : 0 (bf16);
 * db_parts: parts buffer or NULL receiving the partial sums of dz[..., c] (BiasAddGrad of the last conv, fused). */
int abl_gap_bwd_relu(const float* dgap, const void* a, void* dz, float* db_parts, long long db_parts_bytes, int B, int HW,
                     int C, void* stream);
/* the same dz from the 1-bit ReLU mask abl_conv3x3_igemm_gap wrote (relu_bits uint32 [B,HW,C/32]); C % 32 == 0. */
int abl_gap_bwd_bits(const float* dgap, const unsigned int* relu_bits, void* dz, float* db_parts, long long db_parts_bytes,
                     int B, int HW, int C, void* stream);

/* Dropout -> Dense(n_classes) -> softmax (models.py:415-416) [+ categorical_crossentropy with
 * class_weight, models.py:426 / train.py:135-136].
 *   gap fp32 [B,C]; w fp32 [C,n]; b fp32 [n]; n <= 8.
 *   training != 0: keep-mask from the counter hash (seed, step, rank, frame, channel) with
 *     P(drop) = dropout, kept values scaled 1/(1-dropout); mask bytes written to keep_mask [B,C]
 *     (if keep_mask_in != 0 the mask is READ from keep_mask instead -- test hook).
 *   y_onehot fp32 [B,n] or NULL (no loss).  class_weight fp32 [n] or NULL.
 *   loss_sum (+)= sum_i w_i * (logsumexp(z_i) - z_i[y_i]) * inv_batch   (fp32 scalar; loss_accumulate != 0 adds to it --
 *            Keras evaluate over many batches -- else overwrites it)
 *   dlogits[i][j] = w_i * inv_batch * (p_ij - y_ij)                      (if dlogits != NULL)
 *   loss_ws  abl_head_loss_ws_bytes(B) bytes, ZEROED ONCE by the caller (needed when loss_sum and y_onehot are given):
 *            per-block loss sums + an integer ticket; the last block to finish adds them in block order, so the batch
 *            loss is one deterministic sum and a single read-modify-write of *loss_sum (the ticket resets itself).
 *   gap_part (or NULL), gap_parts, gap_hw: the GlobalAveragePooling finish folded in -- the per-rectangle sums that
 *            abl_conv3x3_igemm_gap left (fp32 [B * gap_parts][C]); the kernel then WRITES gap[n][c] = (1 / gap_hw) *
 *            sum_p gap_part[n * gap_parts + p][c] (abl_gap_reduce's arithmetic and order) before using it, and no
 *            abl_gap_reduce launch is needed.  With gap_part == NULL, gap is an input. */
long long abl_head_loss_ws_bytes(int B);
int abl_head_fwd(float* gap, const float* w, const float* b, float* logits, float* probs,
                 const float* y_onehot, const float* class_weight, float inv_batch, float* loss_sum,
                 float* dlogits, unsigned char* keep_mask, int keep_mask_in, int training, float dropout,
                 unsigned long long seed, unsigned int step, unsigned int rank, int B, int C, int n,
                 int loss_accumulate, float* loss_ws, const float* gap_part, int gap_parts, int gap_hw, void* stream);
/* abl_head_fwd with the dropout step counter read from DEVICE memory instead of passed by value, so the launch can be
 * replayed from a CUDA graph (mask always generated; y_onehot and loss_sum required).  step_dev: uint32[2] = {dropout
 * step, optimizer step}; the kernel's last block increments BOTH after every block has read the dropout step, so a
 * captured train step needs no separate counter kernels (abl_adam_flat_tab reads step_dev[1] later in the step). */
int abl_head_fwd_stepdev(float* gap, const float* w, const float* b, float* logits, float* probs,
                         const float* y_onehot, const float* class_weight, float inv_batch, float* loss_sum,
                         float* dlogits, unsigned char* keep_mask, int training, float dropout, unsigned long long seed,
                         unsigned int* step_dev, unsigned int rank, int B, int C, int n, int loss_accumulate,
                         float* loss_ws, const float* gap_part, int gap_parts, int gap_hw, void* stream);
/* dw[C][n] = h^T dlogits, db[n] = sum_i dlogits (both OVERWRITTEN; fixed summation order),
 * dgap[i][c] = scale_ic * sum_j dlogits[i][j] w[c][j] with h = gap * scale, scale_ic = keep_mask ? mask/(1-dropout) : 1. */
int abl_head_bwd(const float* gap, const float* w, const float* dlogits, const unsigned char* keep_mask,
                 float dropout, float* dw, float* db, float* dgap, int B, int C, int n, void* stream);

/* Keras 2.9 Adam (models.py:397): m,v update; theta -= alpha_t * m / (sqrt(v) + eps), where the caller
 * passes alpha_t = lr * sqrt(1-b2^t) / (1-b1^t).  Flat fp32 buffers of n elements. */
int abl_adam_flat(float* theta, const float* g, float* m, float* v, long long n, float alpha_t, float b1,
                  float b2, float eps, void* stream);
/* abl_adam_flat for CUDA-graph replays: alpha_t = alpha_tab[min(*t_dev, tab_len) - 1] with the 1-based step count t in
 * DEVICE memory and a table the host filled with lr*sqrt(1-b2^t)/(1-b1^t) for t = 1..tab_len. */
int abl_adam_flat_tab(float* theta, const float* g, float* m, float* v, long long n, const float* alpha_tab,
                      const int* t_dev, int tab_len, float b1, float b2, float eps, void* stream);
/* Keras 2.9 RMSprop (models.py:398), momentum 0, not centered. */
int abl_rmsprop_flat(float* theta, const float* g, float* rms, long long n, float lr, float rho, float eps,
                     void* stream);

/* tf.image.resize(image, IMG_DIM, method=NEAREST_NEIGHBOR) of the reference's input pipeline (src/data/preprocessor.py:
 * 76-80) on raw uint8 RGB frames: src [B,h,w,3] -> dst [B,H,W,3], TF2 half-pixel-centre rule
 * src_index = min(floor((dst_index + 0.5) * in / out), in - 1).  The output is what abl_conv_first_*_u8 consume. */
int abl_resize_nearest_u8(const unsigned char* src, unsigned char* dst, int B, int h, int w, int H, int W, void* stream);

/* Training-time augmentation + input scaling of one batch: the reference's tf.keras.Sequential([RandomZoom,
 * RandomTranslation, RandomRotation, RandomFlip('horizontal'), RandomBrightness]) (src/data/preprocessor.py:25-32,
 * applied per batch at preprocessor.py:56-57) followed by vgg16.preprocess_input (preprocessor.py:60-63), in one launch.
 *   src     [B,H,W,3] RGB frames at IMG_DIM in [0,255]: uint8 (src_is_u8) or fp32
 *   params  fp32 [B,20] per frame: the output->input affine maps [a0 a1 a2 b0 b1 b2] of the zoom, translation and
 *           rotation layers in Keras' ImageProjectiveTransformV3 convention (x_in = a0*x + a1*y + a2, y_in = b0*x +
 *           b1*y + b2), then flip (0/1) and the brightness delta that is ADDED to every value without clipping (the
 *           reference's RandomBrightness layer, preprocessor.py:83-104, is tf.image.adjust_brightness on a float
 *           image).  The host draws
 *           them (ab_line_classifier_b200/augment.py); the three resamplings are applied one after the other with
 *           bilinear interpolation and constant fill 0, exactly as the three Keras layers do.
 *   dst     fp32 [B,H,W,3]; preprocess != 0: BGR minus [103.939, 116.779, 123.68] (what block1_conv1 consumes),
 *           preprocess == 0: the augmented RGB frame (what the Sequential returns).
 *   workspace: abl_augment_workspace_bytes(B,H,W) bytes, 0 (NULL allowed) when two fp32 planes of a frame fit in one
 *           SM's shared memory (H*W <= 29056, e.g. 128x128). */
long long abl_augment_workspace_bytes(int B, int H, int W);
int abl_augment_frames(const void* src, int src_is_u8, const float* params, float* dst, void* workspace,
                       long long workspace_bytes, int B, int H, int W, int preprocess, void* stream);

/* Device-side state of the Keras metrics the reference compiles the model with (src/train.py:109-115: 'accuracy',
 * AUC(name='auc'), Precision/Recall(thresholds=1/n_classes, class_id=c)), accumulated batch by batch from the one-hot
 * labels and the softmax output so fit()/evaluate() read back abl_metrics_counter_count() integers per epoch instead
 * of every probability.  counters (uint64, zeroed by the caller at epoch start):
 *   [0] frames, [1] frames with argmax(y) == argmax(p), [2+3c ..] tp, fp, fn of class c at class_thresholds[c],
 *   then (num_thresholds+1) buckets for positive and (num_thresholds+1) for negative flattened (y, p) entries: bucket k
 *   counts entries whose p is greater than exactly k of the ascending auc_thresholds (fp32, device), so
 *   true_positives(t) = sum_{k>t} pos[k] as keras.metrics.AUC accumulates it. */
int abl_metrics_counter_count(int n_classes, int num_thresholds);
int abl_metrics_accumulate(const float* probs, const float* y_true, int B, int n_classes, const float* auc_thresholds,
                           int num_thresholds, const float* class_thresholds, unsigned long long* counters,
                           void* stream);

/* Developer probe (not used by the product path): SM clocks per tcgen05.mma (M = 128, K = 16, bf16, operands in shared
 * memory) for a given N and issue pattern -- mode 0 plain, 1 weight-stationary pairs, 2 a different filter window per
 * MMA.  `ctas` CTAs (one per SM) each issue 4*iters MMAs (8*iters in mode 1) and write two uint64 to out_cycles
 * [ctas][2]: clocks of the issue loop and clocks until the last MMA retired.  tools/mma_rate.py prints the table. */
int abl_dev_mma_rate_probe(int n, int mode, int iters, int ctas, unsigned long long* out_cycles, void* stream);
/* Developer probe: fill `bytes` of device memory with SM-issued stores (mode 0 st.v4, 1 st.cs, 2 st.wt, 3 st + L2
 * evict_first, 4 st.v8, 5 TMA bulk stores from shared memory, 6 as 5 with an L2 evict_first hint), ctas_per_sm x #SM CTAs
 * of 512 threads.  tools/write_bw.py times it: the write-side roofline of the HBM-bound kernels. */
int abl_dev_write_probe(void* dst, long long bytes, int mode, int ctas_per_sm, void* stream);

/* Clip-level aggregation of frame probabilities (src/predict.py:159-166 with the helpers at 283-308 and 365-376),
 * all clips in one launch instead of the reference's per-clip Python loop.  probs fp32 [N_frames, n_classes] with the
 * frames of clip c in rows clip_offsets[c] .. clip_offsets[c+1]-1 (clip_offsets: n_clips+1 int64, device).
 *   mean_out[c][j] = mean_f probs[f][j]                                  ('average')
 *   run_out[c]     = longest run of frames with probs[f][b_idx] > threshold   ('contiguous')
 *   win_out[c]     = max over i of mean(probs[i:i+window, b_idx]), 0 if the clip has fewer than `window` frames
 *                                                                        ('sliding_window') */
int abl_clip_aggregate(const float* probs, const long long* clip_offsets, int n_clips, int n_classes, int b_idx,
                       float threshold, int window, float* mean_out, int* run_out, float* win_out, void* stream);

/* Data-parallel optimizer step over NVLink peer memory (new: the reference is single-device).  ONE kernel does
 * reduce-scatter(grad) -> Adam (kind 0) / RMSprop (kind 1) on this rank's 1/world shard -> all-gather(theta):
 *   theta_peers[r], grad_peers[r] (HOST arrays of `world` DEVICE pointers) are the bases of rank r's flat fp32 theta /
 *   grad buffers as mapped into THIS process (a symmetric allocation; entry `rank` is the local buffer); the run
 *   [offset, offset + n) of every buffer is processed (multiples of 4 floats).
 *   g = sum_r grad_r[i] in fixed rank order; theta updated with the local m_or_rms / v state (same offset); the new
 *   value is stored into every rank's theta.  Arguments as abl_adam_flat (lr_or_alpha = alpha_t, b1_or_rho = beta1)
 *   or abl_rmsprop_flat (lr, rho; b2 and v unused).  alpha_tab / t_dev / tab_len (Adam, may be NULL / 0): as in
 *   abl_adam_flat_tab -- alpha_t read on the device, so the step can be replayed from a CUDA graph on every rank.
 * The caller must place a cross-rank barrier before (all gradients final) and after (all weights delivered). */
int abl_opt_flat_dp(int kind, float* const* theta_peers, const float* const* grad_peers, float* m_or_rms, float* v,
                    long long offset, long long n, int rank, int world, float lr_or_alpha, float b1_or_rho, float b2,
                    float eps, const float* alpha_tab, const int* t_dev, int tab_len, void* stream);

/* Grad-CAM of the predicted class for a GAP->Dense->softmax head (src/explainability/gradcam.py:68-81),
 * per frame: alpha_k = (1/HW) sum_j p_c (delta_cj - p_j) w[k][j]; heat = relu(mean_k alpha_k a[h,w,k]);
 * heat /= max(heat) (all-zero map stays zero); bilinear resize (cv2 INTER_LINEAR semantics) to
 * out_h x out_w.  a bf16 [B,H,W,C]; probs fp32 [B,n]; w fp32 [C,n]; heat fp32 [B,out_h,out_w]. */
int abl_gradcam(const void* a, const float* probs, const float* w, float* heat, int B, int H, int W, int C,
                int n, int out_h, int out_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ABL_B200_H_ */
