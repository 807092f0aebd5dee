/*
 * cffm_b200.h -- C ABI of the B200-native CFFM hot path (libcffm_b200.so).
 *
 * Drop-in boundary.  The reference (GuoleiSun/VSS-CFFM) is pure Python on PyTorch + mmcv
 * (setup.py:125 ext_modules=[]): it has no FFI of its own.  Its hot path is the chain of
 * nn.Module.forward methods below; a maintainer binds this library with ctypes (see
 * INTEGRATION.md) from inside those forward methods.  Each entry point names the reference
 * code it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers and sizes; no torch types.  Every pointer is a DEVICE pointer on the
 *    current CUDA device unless it is documented as host memory.
 *  - activations are token-major ("NHWC"): row = (frame, y, x), channels contiguous.
 *    `half` = IEEE fp16 (GEMM / attention operands), `float` = fp32 (residual streams,
 *    LayerNorm / softmax statistics, accumulators).  ld* = row stride in ELEMENTS.
 *  - weights are fp16 [N_out, K] row-major (torch.nn.Linear layout).
 *  - `stream` is a cudaStream_t passed as void*.  All calls are asynchronous on it, never
 *    allocate, never synchronise, keep no global state (re-entrant per stream; capturable
 *    in a CUDA graph).  The caller owns all memory (PyTorch stays the allocator).
 *  - return value: 0 = ok, >0 = CFFM_E_* (bad argument / unsupported shape / arch),
 *    <0 = -(cudaError_t).  Nothing throws across the ABI.  cffm_last_error() gives a
 *    thread-local human-readable message for the last non-zero status.
 *  - there is NO CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef CFFM_B200_H_
#define CFFM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFFM_ABI_VERSION 1

enum {
  CFFM_OK = 0,
  CFFM_E_BADARG = 1,      /* null pointer, negative size, misaligned pointer or stride   */
  CFFM_E_UNSUPPORTED = 2, /* shape outside what the kernels are built for                */
  CFFM_E_ARCH = 3,        /* current device is not compute capability 10.x               */
  CFFM_E_DRIVER = 4       /* cuTensorMapEncodeTiled unavailable / failed                 */
};

enum { CFFM_ACT_NONE = 0, CFFM_ACT_GELU = 1, CFFM_ACT_RELU = 2 };
enum { CFFM_GEMM_TCGEN05 = 0, CFFM_GEMM_CHECK = 1 };

int cffm_abi_version(void);
const char* cffm_last_error(void);
/* 0 when the current device can run this library (sm_100 family). */
int cffm_device_check(void);
/* Ordinal of the device the library will launch on (the caller's current CUDA context), or
 * -(cudaError_t).  Lets a multi-process host assert that rank r really drives GPU r. */
int cffm_current_device(void);

/* out = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N]); fp16 operands, fp32 accumulate.
 * Replaces every nn.Linear / 1x1 nn.Conv2d / (with cffm_im2col) k x k nn.Conv2d on the path:
 * mix_transformer.py:49,53,98,104,114,173-195  cffm_head.py:36,119,123,148
 * cffm_transformer.py:21-24,374,449,495,602  swin_transformer_2d.py:219,223,258.
 * bias, residual, out_f16, out_f32 may each be NULL (at least one output is required).
 * residual is fp32 [M, ldr] and may alias out_f32 (in-place residual stream).
 * Requires K % 8 == 0, N % 8 == 0, lda/ldw % 8 == 0, 16-byte aligned A/W/outputs.
 * impl: CFFM_GEMM_TCGEN05 = TMA + tcgen05.mma + TMEM kernel (the product path);
 *       CFFM_GEMM_CHECK   = simple CUDA-core kernel used by the tests to cross-check. */
int cffm_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                  const float* residual, int64_t ldr, void* out_f16, int64_t ldo16,
                  float* out_f32, int64_t ldo32, int M, int N, int K, int act, int impl,
                  void* stream);

/* GEMM with a LayerNorm fused into its epilogue, for N <= 128 (one tile spans the row):
 *   x = A.W^T + bias (+ residual)  -> out_f32 (may be NULL; may alias residual);
 *   LayerNorm(x; ln_gamma, ln_beta, ln_eps) -> ln_out_f16 (fp16 [M, ldln]).
 * The MiT pairs "proj / fc2 + residual" followed by norm2 / next norm1 / the stage norm
 * (mix_transformer.py:154-155,321-343) as one kernel instead of two. */
int cffm_gemm_f16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                     const float* residual, int64_t ldr, float* out_f32, int64_t ldo32,
                     const float* ln_gamma, const float* ln_beta, float ln_eps, void* ln_out_f16,
                     int64_t ldln, int M, int N, int K, void* stream);

/* The same with TWO chained LayerNorms (OverlapPatchEmbed.norm, mix_transformer.py:198, followed by the first block's
 * norm1, :154) and no residual: x = A.W^T + bias;  y = LayerNorm(x; gamma1, beta1, eps1) -> out_f32 [M,N];
 * LayerNorm(y; gamma2, beta2, eps2) -> ln_out_f16 [M,N].  N <= 128. */
int cffm_gemm_f16_ln_chain(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* out_f32,
                           int64_t ldo32, const float* gamma1, const float* beta1, float eps1, const float* gamma2,
                           const float* beta2, float eps2, void* ln_out_f16, int64_t ldln, int M, int N, int K,
                           void* stream);

/* Split-K GEMM for the few-tile / long-K convolutions of the path (spatial-reduction conv
 * mix_transformer.py:76,101-102 with K = sr*sr*C up to 4096; patch embeds of stages 3-4 with K = 9*C):
 * split s covers a contiguous range of 64-wide k-blocks and writes its fp32 partial product to
 * partials[s] ([splits, M, N] contiguous, no bias); cffm_layernorm_sum adds them in a fixed order
 * (deterministic).  cffm_splitk_plan returns the split count to use (1 = do not split); it depends on K
 * only, so that a row's summation order -- hence a clip's result -- does not change with the batch size. */
int cffm_gemm_f16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* partials,
                         int M, int N, int K, int splits, void* stream);
int cffm_splitk_plan(int M, int N, int K);

/* Convolutions as implicit GEMMs on the same tcgen05 kernel (no im2col matrix is written): x fp16 NHWC
 * [n, H, W, C] (C % 64 == 0), filter ksize x ksize, weights fp16 [Nout, ksize*ksize*C] in (ky, kx, c)
 * order.  The A operand is the image itself, read through a 4-D tensor map whose boxes walk it with the
 * convolution stride (one box per filter tap and 64-channel chunk); out-of-image taps are the TMA's zero
 * fill = the convolution's zero padding.  Needs Wo <= 128 and Wo * stride <= 256 (one box per output
 * row group); larger images go through cffm_im2col + cffm_gemm_f16.
 * OverlapPatchEmbed.proj (k3 s2 p1) and Attention.sr (k = stride), mix_transformer.py:76,101-102,173-195.
 *   _ln     : y = conv + bias -> out_f32 (optional) ; LayerNorm(y; gamma1, beta1) -> ln_out_f16; with gamma2 != NULL the
 *             first LayerNorm's result goes to out_f32 and a second LayerNorm (gamma2, beta2) of it to ln_out_f16
 *             (patch-embed norm + first block's norm1).  Nout <= 128.
 *   _splitk : partials[s] = the s-th K range of the convolution, fp32 [splits, n*Ho*Wo, Nout] (summed by
 *             cffm_layernorm_sum / cffm_layernorm_chain). */
int cffm_conv_gemm_f16_ln(const void* x, int n, int H, int W, int C, int ksize, int stride, int pad, const void* Wt,
                          int64_t ldw, const float* bias, float* out_f32, int64_t ldo32, const float* gamma1,
                          const float* beta1, float eps1, const float* gamma2, const float* beta2, float eps2,
                          void* ln_out_f16, int64_t ldln, int Nout, void* stream);
int cffm_conv_gemm_f16_splitk(const void* x, int n, int H, int W, int C, int ksize, int stride, int pad,
                              const void* Wt, int64_t ldw, float* partials, int Nout, int splits, void* stream);

/* Stage-1 OverlapPatchEmbed in ONE kernel (no patch matrix): y = LayerNorm(conv7x7_s4_p3(x) + bias; gamma1, beta1) -> out_f32
 * [n*Ho*Wo, Nout] (the fp32 residual stream), LayerNorm(y; gamma2, beta2) -> ln_out_f16 (the first block's norm1).
 * OverlapPatchEmbed.forward mix_transformer.py:173-200 followed by Block.norm1 :154.
 *   x  fp32 [n, 3, H, W] (NCHW, the normalised frames), W %% 4 == 0;
 *   Wk fp16 [Nout, 192]: Wk[o, (c*7 + ky)*8 + kx] = conv.weight[o, c, ky, kx], zero elsewhere (kx = 7, columns >= 168).
 * cffm_patch_embed_s1_supported(...) != 0 names what it is built for (3 channels, k7 s4 p3, Nout in {32, 64}); other
 * geometries: cffm_im2col + cffm_gemm_f16_ln_chain. */
int cffm_patch_embed_s1_supported(int W, int Cin, int ksize, int stride, int pad, int Nout);
int cffm_patch_embed_s1(const float* x, int n, int H, int W, const void* Wk, const float* bias, const float* gamma1,
                        const float* beta1, float eps1, const float* gamma2, const float* beta2, float eps2,
                        float* out_f32, void* ln_out_f16, int Nout, void* stream);

/* Mix-FFN tail in ONE kernel: x_out = residual + fc2(GELU(dwconv3x3(hidden) + dw_b)) + b2, and optionally
 * LayerNorm(x_out; gamma, beta, eps) -> ln_out_f16.  Replaces Mlp.dwconv + act + fc2 (mix_transformer.py:52-58, DWConv
 * :361-368), the residual add and the next norm (Block.forward :84-88): the convolved / activated hidden map stays on chip.
 *   hidden fp16 [n, H, W, HD] (NHWC) = fc1 output incl. bias; dw_w fp16 [9, HD] (tap-major: ky*3+kx), dw_b fp32 [HD];
 *   W2 fp16 [N, HD] (row stride ldw2), b2 fp32 [N]; residual fp32 [n*H*W, N]; out_f32 fp32 [n*H*W, N] or NULL (may
 *   alias residual); ln_out_f16 fp16 [n*H*W, N] or NULL (then gamma / beta are ignored).
 * cffm_mixffn_tail_supported(N, HD) != 0 names the shapes it is built for (N in {64, 128}, HD % 64 == 0, HD <= 512);
 * other shapes: cffm_dwconv3x3_gelu + cffm_gemm_f16(_ln). */
int cffm_mixffn_tail_supported(int N, int HD);
int cffm_mixffn_tail(const void* hidden, int n, int H, int W, int HD, const void* dw_w, const float* dw_b, const void* W2,
                     int64_t ldw2, const float* b2, const float* residual, float* out_f32, const float* ln_gamma,
                     const float* ln_beta, float ln_eps, void* ln_out_f16, int N, void* stream);

/* LayerNorm of x = sum_s partials[s] + bias (partials fp32 [nsum, M, C] contiguous; bias may be NULL):
 * the reduction of cffm_gemm_f16_splitk fused into the LayerNorm that follows every such conv
 * (mix_transformer.py:103,198). */
int cffm_layernorm_sum(const float* partials, int nsum, const float* bias, const float* gamma,
                       const float* beta, float eps, void* out_f16, int64_t ldo16, float* out_f32,
                       int64_t ldo32, int M, int C, void* stream);

/* Two chained LayerNorms in one pass: y = LayerNorm(sum_s partials[s] + bias; gamma, beta, eps) -> out_f32 [M,C];
 * LayerNorm(y; gamma2, beta2, eps2) -> out_f16 [M,C].  OverlapPatchEmbed.norm followed by the first block's norm1
 * (mix_transformer.py:198 then :154).  partials fp32 [nsum, M, C] contiguous; bias may be NULL. */
int cffm_layernorm_chain(const float* partials, int nsum, const float* bias, const float* gamma, const float* beta,
                         float eps, float* out_f32, int64_t ldo32, const float* gamma2, const float* beta2,
                         float eps2, void* out_f16, int64_t ldo16, int M, int C, void* stream);

/* Row LayerNorm over C channels, fp32 statistics.  x is fp32 (x_is_f32=1) or fp16.
 * Writes fp16 and/or fp32.  mix_transformer.py:154-155,198,321  cffm_transformer.py:824
 * swin_transformer_2d.py:619,622,663. */
int cffm_layernorm(const void* x, int x_is_f32, int64_t ldx, const float* gamma, const float* beta,
                   float eps, void* out_f16, int64_t ldo16, float* out_f32, int64_t ldo32,
                   int M, int C, void* stream);

/* Patch extraction for conv-as-GEMM: A[(n,oy,ox), (ky,kx,c)] fp16 with row stride Kpad
 * (columns >= k*k*C are zero-filled).  layout 0: x = fp32 NCHW image (N,C,H,W);
 * layout 1: x = fp16 NHWC (N,H,W,C).  Zero padding `pad`.
 * OverlapPatchEmbed.proj (mix_transformer.py:173-195) and the k=s spatial-reduction conv
 * Attention.sr (mix_transformer.py:76,101-102). */
int cffm_im2col(const void* x, int layout, int N, int H, int W, int C, int k, int stride, int pad,
                void* A, int Kpad, void* stream);

/* softmax(scale * q k^T) v for small key sets held in shared memory.
 * q [batch, Nq, heads*head_dim] (row stride ldq), k/v [batch, Nkv, ...] (row stride ldkv),
 * out [batch, Nq, heads*head_dim] (row stride ldo); head h uses columns [h*head_dim, ...).
 * head_dim in {32, 64}.  MiT efficient attention (mix_transformer.py:109-113) and the CFFM++
 * prototype cross-attention (swin_transformer_2d.py:226-257). */
int cffm_mha_f16(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                 int64_t ldo, int batch, int Nq, int Nkv, int heads, int head_dim, float scale,
                 void* stream);

/* Mix-FFN middle: depthwise 3x3 conv (pad 1) + bias + exact GELU on fp16 NHWC.
 * w is fp16 [9, C] (tap-major), bias fp32 [C].  mix_transformer.py:50-51,361-368. */
int cffm_dwconv3x3_gelu(const void* x, const void* w, const float* bias, void* out, int N, int H,
                        int W, int C, void* stream);

/* SegFormer MLP-decoder fuse after weight folding (linear_c{i} . linear_fuse . BN folded on the
 * host): c = relu(p1 + up(p2) + up(p3) + up(p4) + shift), bilinear align_corners=False
 * upsampling of the coarser maps to (H1,W1).  p_i fp16 NHWC [N,Hi,Wi,C].
 * c_full (fp16 [N,H1,W1,C], may be NULL) receives _c; c_half_f32 / c_half_f16 (may be NULL)
 * receive the 2x2 mean = resize(_c, 1/2) (cffm_head.py:108-119,131-133).
 * T_perm > 1: the N input frames are in the reference's clip-major order (n = b*T + t,
 * encoder_decoder.py:557) and every output is written frame-major (slot t*B + b), the layout
 * the CFFA/CFM entry points use (target frames contiguous at the end); T_perm <= 1: no reorder. */
int cffm_head_fuse(const void* p1, const void* p2, const void* p3, const void* p4, int N, int H1,
                   int W1, int H2, int W2, int H3, int W3, int H4, int W4, int C, int T_perm,
                   const float* shift, void* c_full, float* c_half_f32, int64_t ldh32,
                   void* c_half_f16, int64_t ldh16, void* stream);

/* CFFA step 1: LayerNorm (norm1) of all T frames of x fp32 [T,B,H,W,C] (frame-major; the target
 * frames are the last B) -> xn fp16 (same shape)
 * and the target map in the layout the CFM attention reads: xt_pad fp16 [B,Hp+6,Wp+6,C], the
 * zero-padded LN map (pad AFTER norm) with a 3-wide CYCLIC apron: position (Y,X) holds the padded
 * map at ((Y-3) mod Hp, (X-3) mod Wp), which materialises the torch.roll wrap-around of the
 * neighbour windows.  Every position (pad zeros included) is written on every call.
 * cffm_transformer.py:713-734, :389-400. */
int cffm_cffa_norm(const float* x, const float* gamma, const float* beta, float eps, void* xn,
                   void* xt_pad, int B, int T, int H, int W, int Hp, int Wp, int C, void* stream);

/* Same kernel on an arbitrary list of frames (frame-sharded multi-GPU path): x fp32 [n_frames,H,W,C];
 * frames >= first_target are target frames and are also written, in the apron layout above, to xt_pad
 * [n_frames-first_target,Hp+6,Wp+6,C] (xt_pad may be NULL when first_target == n_frames). */
int cffm_cffa_norm_frames(const float* x, const float* gamma, const float* beta, float eps, void* xn,
                          void* xt_pad, int n_frames, int first_target, int H, int W, int Hp, int Wp,
                          int C, void* stream);

/* CFFA step 2: coarse-to-fine pooling of the (virtually zero-padded) LN'ed frames.
 * xn fp16 [T,B,H,W,C] frame-major.  pooled fp16 [B, P, C], P = nW*(1+1+4+9): target 7x7 fc-pool | ref0 7x7 | ref1 bilinear
 * (Hp->6*nWh) + 3x3 | ref2 bilinear + 2x2, each map row-major.  pool_w fp32 packed
 * [49+49+9+4], pool_b fp32 [4].  cffm_transformer.py:739-805. */
int cffm_cffa_pool(const void* xn, int B, int T, int H, int W, int C, const float* pool_w,
                   const float* pool_b, void* pooled, void* stream);

/* The same pooling split in two, writing DISJOINT rows of the same pooled [B, P, C] buffer, so that the
 * reference-frame levels (which never change while the blocks update the target) can be produced ahead of
 * time on another stream.  part 0: the target level only, xn fp16 [B,H,W,C] = the LN'ed target frames;
 * part 1: the three reference levels only, xn fp16 [3,B,H,W,C] = the LN'ed reference frames, frame-major. */
int cffm_cffa_pool_part(const void* xn, int B, int part, int H, int W, int C, const float* pool_w,
                        const float* pool_b, void* pooled, void* stream);

/* One pooling level of n independent LN'ed frames xn fp16 [n,H,W,C] (frame-sharded multi-GPU path: the
 * owner of a reference frame pools it for its temporal role).  level 0: target 7x7 | 1: ref0 7x7 |
 * 2: ref1 bilinear + 3x3 | 3: ref2 bilinear + 2x2; pooled fp16 [n, {1,1,4,9}[level]*nW, C]. */
int cffm_cffa_pool_level(const void* xn, int n_frames, int level, int H, int W, int C,
                         const float* pool_w, const float* pool_b, void* pooled, void* stream);

/* Cross-frame feature mining attention (TMA-assembled K/V, tcgen05 QK^T / PV; no roll / partition /
 * unfold / cat is materialised).
 *   qkv_a     fp16 [B, Hp+6, Wp+6, 3C] = qkv(xt_pad) in the apron layout of cffm_cffa_norm
 *   kv_pooled fp16 [B, 15 nW, 2C]      = K,V thirds of qkv(pooled); level maps target 9x9-like | ref0 |
 *                                        ref1 | ref2 at token offsets {0, 1, 2, 6} nW
 *   bias_tab  fp16 [heads, 49, pitch]  = window-independent additive logit term DIVIDED by `scale`, in
 *                                        the kernel's key order (cffm_cfm_layout): 13x13 halo of the
 *                                        window row-major (own window + ring; a ring key the reference
 *                                        lists twice carries logaddexp of its two entries), then the
 *                                        5x5 / 7x7 / 5x5 / 3x3 pooled windows; -inf on unused columns
 *   out       fp16 [B, H*W, C]: window_reverse + crop already applied.
 * C = 256, heads = 8, window 7, expand 3 (the head's hard-coded setting, cffm_head.py:74-95).
 * cffm_transformer.py:364-601 and :809-821. */
int cffm_cfm_attention(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, int B,
                       int H, int W, int C, int heads, float scale, void* stream);

/* Same kernel with the pooled K/V addressed level by level (frame-sharded multi-GPU runs read the
 * all-gathered reference-frame buffer in place, no per-clip copies):
 *   kv_tgt       pooled target level fp16 [B, nW, 2C], clip stride tgt_stride (elements)
 *   kv_ref{k}    base of the level maps of reference role k (nW / 4 nW / 9 nW tokens x 2C per map); the map of
 *                slot s contributed by rank r starts s * slot_stride{k} + r * rank_stride elements further
 *   n_slots{k}   slots per rank of role k; n_ranks = ranks behind the buffer
 *   ref_slot     DEVICE table int32 [B, 3, 2]: (slot, rank) of reference frame k of clip b; NULL = (b, 0) */
int cffm_cfm_attention_slots(const void* qkv_a, const void* kv_tgt, int64_t tgt_stride, const void* kv_ref0,
                             const void* kv_ref1, const void* kv_ref2, int64_t slot_stride0,
                             int64_t slot_stride1, int64_t slot_stride2, int n_slots0, int n_slots1,
                             int n_slots2, int64_t rank_stride, int n_ranks, const int32_t* ref_slot,
                             const void* bias_tab, void* out, int B, int H, int W, int C, int heads, float scale,
                             void* stream);

/* Test entry: the same kernel, additionally writing the assembled K and V tiles of every work item
 * (head pair hp, clip b, window w) to dump fp16 [4, B*nW, 2, NPAD, 64] (K then V; row order =
 * cffm_cfm_layout) so that the TMA assembling can be compared bit for bit with the reference's
 * roll / window_partition / valid_ind_rolled / nn.Unfold / cat pipeline (cffm_transformer.py:378-522). */
int cffm_cfm_attention_dump(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out,
                            void* dump, int B, int H, int W, int C, int heads, float scale, void* stream);

/* Key-row layout of the CFM kernel (host only, no GPU work): out8 = {row of the halo block (0), of the
 * pooled target block, of reference 0, 1, 2, number of key rows NPAD, bias_tab row pitch, apron width}. */
int cffm_cfm_layout(int32_t* out8);

/* Bilinear (align_corners=False) resize of NHWC logits to fp32 NCHW.
 * in: fp16 (in_is_f32=0) or fp32 [B,h,w,ldc] using the first ncls channels.
 * cffm_head.py:149 (x2 -> 1/4 scale), mmseg/ops/wrappers.py:8-29. */
int cffm_resize_nhwc_to_nchw(const void* in, int in_is_f32, int64_t ldc, float* out, int B, int h,
                             int w, int ncls, int Ho, int Wo, void* stream);

/* Fused bilinear resize (align_corners=False) of fp32 NCHW logits to (Ho,Wo) + argmax over
 * channels -> int64 labels [B,Ho,Wo].  softmax is monotone, so it is skipped.
 * encoder_decoder.py:373-377,542,564. */
int cffm_resize_argmax(const float* logits, int64_t* labels, int B, int ncls, int h, int w, int Ho,
                       int Wo, void* stream);

/* cffm_head.py:149 + encoder_decoder.py:373-377,542,564 in one pass: NHWC fp32 class scores
 * [B,h,w,ldc] (first ncls channels) -> bilinear to (Hm,Wm) -> bilinear to (Ho,Wo) -> argmax ->
 * int64 labels [B,Ho,Wo]; the two intermediate logit tensors are never written.  Requires the second
 * stage to upsample by >= ~3.2x (else CFFM_E_UNSUPPORTED: use cffm_resize_nhwc_to_nchw +
 * cffm_resize_argmax). */
int cffm_upsample2_argmax(const float* scores, int64_t ldc, int64_t* labels, int B, int h, int w,
                          int ncls, int Hm, int Wm, int Ho, int Wo, void* stream);
/* The same with 8-bit labels (ncls <= 256): an eighth of the bytes that leave the device per step. */
int cffm_upsample2_argmax_u8(const float* scores, int64_t ldc, uint8_t* labels, int B, int h, int w,
                             int ncls, int Hm, int Wm, int Ho, int Wo, void* stream);

/* Bilinear (align_corners=False) resize of fp32 NCHW maps [B,C,h,w] -> [B,C,Ho,Wo]: the extra
 * "rescale to ori_shape" step of whole_inference (encoder_decoder.py:507-514). */
int cffm_resize_nchw(const float* in, float* out, int B, int C, int h, int w, int Ho, int Wo,
                     void* stream);

/* softmax over the class dimension of fp32 NCHW logits (encoder_decoder.py:542); only needed when
 * a caller asks for probabilities -- simple_test's argmax skips it. */
int cffm_softmax_nchw(const float* in, float* out, int B, int C, int64_t HW, void* stream);

/* ---- k-means prototypes of CFFM++ (cffm_head.py:267-294: fast_pytorch_kmeans.KMeans(n_clusters, max_iter=10,
 * mode='euclidean').fit_predict on the 1/8-scale decoder features of one video).  One Lloyd iteration is
 *   scores = X c_hi^T + X c_lo^T              two cffm_gemm_f16 calls (fp32 out, second one accumulating)
 *   cffm_kmeans_assign                         labels, one-hot [Kp, Np_pad] fp16, counts (caller zeroes counts first)
 *   partials = onehot . Xt^T                   cffm_gemm_f16_splitk
 *   cffm_kmeans_update                         centroids, error, hi/lo halves and norms for the next iteration
 * Kp = K rounded up to the GEMM tile (128), Np_pad = Np rounded up to 8. */

/* fp32 centroids [K,E] -> fp16 hi / lo halves [Kp,E] (rows >= K zero) and squared norms [Kp]. */
int cffm_kmeans_prepare(const float* centroids, void* c_hi, void* c_lo, float* cnorm, int K, int Kp, int E,
                        void* stream);

/* labels[p] = arg max_j (2 scores[p][j] - cnorm[j]) (first maximum), int64 [Np]; onehot fp16 [Kp, Np_pad];
 * counts int32 [K] += members (atomic).  scores fp32 [Np, lds]. */
int cffm_kmeans_assign(const float* scores, int64_t lds, const float* cnorm, int Np, int Np_pad, int K, int Kp,
                       int64_t* labels, void* onehot, int* counts, void* stream);

/* centroids[j] <- sum_s partials[s][j] / counts[j] (0 for an empty cluster, as the library zeroes the NaN);
 * *error = sum_j |new - old|^2 (summed in cluster order by the last CTA; `done` is a zero-initialised counter
 * the kernel resets).  partials fp32 [nsplit, Kp, E]. */
int cffm_kmeans_update(const float* partials, int nsplit, const int* counts, float* centroids, void* c_hi, void* c_lo,
                       float* cnorm, float* err_per_cluster, float* error, unsigned int* done, int K, int Kp, int E,
                       void* stream);

/* x fp16 [rows, cols] -> xt fp16 [cols, rows_pad] (columns >= rows zero). */
int cffm_transpose_f16(const void* x, int rows, int cols, void* xt, int rows_pad, void* stream);

/* ---- test-time clip preprocessing (mmseg/datasets/pipelines/transforms.py:382-421 AlignedResize_clips, :1277-1297
 * Normalize_clips; the arithmetic is mmcv 1.3.0 -> OpenCV).  Bit-exact with cv2 / mmcv. */

/* cv2.resize(..., INTER_LINEAR) of N uint8 HWC 3-channel images [N,h,w,3] -> [N,H,W,3] (8-bit fixed-point path). */
int cffm_resize_u8(const void* src, int N, int h, int w, void* dst, int H, int W, void* stream);

/* The same resize (the identity when h == H and w == W) fused with mmcv.imnormalize and HWC -> CHW:
 * out[n] (at out + n*out_stride floats) = fp32 [3,H,W], channel c = (pix[to_rgb ? 2-c : c] - mean3[c]) / std3[c]
 * evaluated as float(double(float(x) - mean) * (1 / double(std))).  mean3 / std3 are HOST pointers. */
int cffm_resize_normalize_u8(const void* src, int N, int h, int w, float* out, int64_t out_stride, int H, int W,
                             const float* mean3, const float* std3, int to_rgb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFFM_B200_H_ */
