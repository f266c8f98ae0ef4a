/* autoprog_b200 C ABI — the drop-in boundary of the B200-native AutoProg hot path.
 *
 * The reference (changlin31/AutoProg) has no FFI of its own: its hot path is `nn.Module.forward` calling
 * ATen ops.  Each entry point below is what a Python binding for that path binds instead
 * of the ATen call chain it names (file:line in /root/reference).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C: device pointers + sizes, `apb_stream_t` is a `cudaStream_t` passed as `void*`;
 *   - every tensor is contiguous in the documented layout; the CALLER allocates outputs and workspaces
 *     (PyTorch's caching allocator in the shipped binding); kernels never allocate, free or synchronise;
 *   - dtype codes: APB_F32 = 0, APB_BF16 = 1 (activations); parameters / statistics / loss are fp32;
 *   - return value: 0 = ok, > 0 = cudaError_t of the failed launch, < 0 = argument check
 *     (-1 arg, -2 dtype, -3 shape, -4 unsupported); `apb_last_error()` has the message (thread-local);
 *   - thread-safe: no mutable global state except a mutex-guarded TMA-descriptor cache, the launch / fallback evidence
 *     counters and the diagnostic switches at the end of this header (plain ints, set only by tools, never read from
 *     the environment); all work is enqueued on the given stream (CUDA-graph capturable).
 */
#ifndef AUTOPROG_B200_H_
#define AUTOPROG_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* apb_stream_t;

#define APB_F32 0
#define APB_BF16 1

const char* apb_last_error(void);
int apb_abi_version(void);

/* ---- OutlookAttention core: nn.Unfold -> softmax(scale*logits) -> attn@v -> F.fold  (models/volo.py:83-98)
 * v,y,dy,dv [B,H,W,heads*32] NHWC; logits,dlogits [B,ceil(H/2),ceil(W/2),heads*81] (pre-scale output of the
 * `attn` Linear, channel = head*81 + P*9 + Q).  kernel 3, padding 1, stride 2, head_dim 32 (every VOLO variant).
 * `_simt` = fp32-exact CUDA-core path (parity mode); the un-suffixed entry picks the tensor-core bf16 kernel
 * for APB_BF16 and the SIMT kernel for APB_F32. */
int apb_outlook_fwd_simt(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale,
                         int lpitch, int dtype, apb_stream_t stream);
int apb_outlook_bwd_simt(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H,
                         int W, int heads, float scale, int lpitch, int dtype, apb_stream_t stream);
int apb_outlook_fwd(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                    int dtype, apb_stream_t stream);
int apb_outlook_bwd(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                    int heads, float scale, int lpitch, int dtype, apb_stream_t stream);
/* the two bf16 forward kernels behind apb_outlook_fwd: `_fma` = gather formulation on the CUDA cores (outlook_fma.cu, one
 * warp per 2x2 output block and head pair), `_mma` = mma.sync fragments + staged fold (outlook_mma.cu); both return
 * APB_ERR_UNSUPPORTED when their tile does not fit shared memory. */
int apb_outlook_fwd_fma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                        apb_stream_t stream);
int apb_outlook_fwd_mma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                        apb_stream_t stream);
/* the two bf16 backward kernels behind apb_outlook_bwd: `_fma` = one kernel, dV as the forward's gather with transposed
 * weights + dlogits as a 16x16x32 mma.sync product per (window, head) (outlook_bwd_fma.cu); `_mma` = the band kernel of
 * outlook_mma.cu.  Both return APB_ERR_UNSUPPORTED when their tile does not fit shared memory. */
int apb_outlook_bwd_fma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                        int heads, float scale, int lpitch, apb_stream_t stream);
int apb_outlook_bwd_mma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                        int heads, float scale, int lpitch, apb_stream_t stream);
/* lpitch: elements between consecutive windows in logits/dlogits, heads*81 <= lpitch < heads*81+8.  The bf16 path
 * pads 486 -> 488 so the producing / consuming GEMMs satisfy TMA's 16-byte row pitch; backward zero-fills the pad. */

/* ---- TokenLabelCrossEntropy forward + gradient in one pass  (loss/cross_entropy.py:136-156, :30-36)
 * x_cls [B,C], x_aux [B,N,C] (dtype); target fp32 [B,C,2+N] (target_is_3d=1) or [B,C] (0);
 * box_area = (bbx2-bbx1)*(bby2-bby1); loss: 1 float; d_cls/d_aux: gradients for upstream gradient 1.
 * workspace: apb_tlce_workspace_floats(B,N) floats.
 * ticket: optional DEVICE int32, ZERO on entry and private to the stream (the kernel leaves it zero): with it the
 * class-major fast path is ONE launch -- dense tiles, class-token rows and a last-CTA reduction of the per-CTA loss
 * partials in a fixed order; NULL = three launches (dense, class tokens, reduce).  Same bits either way. */
long long apb_tlce_workspace_floats(int B, int N);
int apb_tlce_fwd_bwd(const void* x_cls, const void* x_aux, const float* target, int target_is_3d, int B, int N, int C,
                     int box_area, const int* box_dev, float w_cls, float w_dense, float* loss, void* d_cls, void* d_aux,
                     float* workspace, int* ticket, int dtype, apb_stream_t stream);
/* box_dev: optional DEVICE int[4] (bbx1,bby1,bbx2,bby2) read by the kernel instead of box_area, so a captured CUDA
 * graph sees a fresh mix-token box on every replay. */
int apb_scale_by_scalar(const void* in, void* out, long long n, const float* scalar, int dtype, apb_stream_t stream);
/* backward of the fused loss: the gradients above were produced for upstream gradient 1; this folds the real upstream
 * gradient g (DEVICE scalar) into both buffers IN PLACE and returns without touching memory when g / applied == 1 (every
 * training step).  applied: device float, initialised to 1 by the caller, remembers the factor already folded in. */
int apb_scale_lazy(void* buf_a, long long n_a, void* buf_b, long long n_b, const float* g, float* applied, int dtype,
                   apb_stream_t stream);

/* ---- token-label TARGET builder  (tlt.data.create_token_label_target at main_prog.py:983-1004, 1919-1932; tlt is
 * un-vendored: recipe restated in oracle/token_label_cpu.py, parity unpinned upstream)
 * maps fp32 [B,3,5,Hm,Wm]: plane 0 top-5 scores, plane 1 top-5 class ids, plane 2 at [0,0,0:6] = crop box x1,y1,x2,y2
 * (normalised), flip flag, ground-truth class.  out fp32 [B,C,2+L*L] = what apb_tlce_fwd_bwd takes as its 3-D target:
 * slot 0 smoothed ground truth, slot 1 class-level label (RoIAlign to 1x1), slots 2.. RoIAlign to LxL; softmax over
 * classes (apply_softmax), then value*on+off.  apb_onehot_smooth: int64 labels [B] -> smoothed one-hot [B,C]. */
int apb_token_label_target(const float* maps, float* out, int B, int C, int Hm, int Wm, int label_size, float smoothing,
                           int apply_softmax, apb_stream_t stream);
int apb_onehot_smooth(const long long* labels, float* out, int B, int C, float smoothing, apb_stream_t stream);

/* ---- fused residual add (+DropPath per-sample scale) + LayerNorm  (models/volo.py:142-143, 232-233, 306-307)
 * xs = x + rs[row / rows_per_sample] * r (r, rs, xs_out optional);  y = LN(xs) * gamma + beta (y optional).
 * bwd: dxs = dres + LN'(dy) (dres optional); dr = rs[b]*dxs (optional, compute dtype: gradient of the branch r).
 * sdtype: dtype of x/xs_out/dres/dxs (residual stream); cdtype: dtype of r/y/dy/dr. gamma,beta,mean,rstd fp32. */
int apb_ln_fwd(const void* x, const void* r, const float* rs, int rows_per_sample, const float* gamma, const float* beta,
               void* xs_out, void* y, float* mean, float* rstd, long long rows, int C, float eps, int sdtype, int cdtype,
               apb_stream_t stream);
long long apb_ln_bwd_workspace_floats(int C);
int apb_ln_bwd(const void* dy, const void* xs, const float* mean, const float* rstd, const float* gamma, const void* dres,
               void* dxs, void* dr, const float* rs, int rows_per_sample, float* dgamma, float* dbeta, int accumulate,
               float* workspace, long long rows, int C, int sdtype, int cdtype, apb_stream_t stream);
long long apb_colsum_workspace_floats(long long rows, int C);
int apb_colsum(const void* a, long long rows, int C, float* out, int accumulate, float* workspace, int dtype,
               apb_stream_t stream);

/* ---- GEMM  C[M,N] = epi( sum_k A(m,k) * B(n,k) (+ bias[n]) )   (nn.Linear / patchify convs: models/volo.py:80,88,100,
 * 161-167,188,199,253-254,370-373,389; fwd = NT, dgrad = NN, wgrad = TN via the transpose flags)
 *   trans_a = 0: A stored [M,K] row-major; 1: A stored [K,M] row-major.
 *   trans_b = 0: B stored [N,K] row-major (nn.Linear weight); 1: B stored [K,N] row-major.
 *   epilogue: 0 none | 1 GELU: u = acc+bias, C = gelu(u), aux = gelu'(u) (out dtype; saved for the backward)
 *             | 2 dGELU: C = acc * aux  | 3 accumulate: C += acc (SIMT only)
 *   in_dtype: dtype of A and B; out_dtype: dtype of C/aux; bias fp32 or NULL.
 * apb_gemm_simt: CUDA-core fp32-accumulate path (exact fp32 products; parity mode, any shape).
 * apb_gemm_tc  : tcgen05/TMEM/TMA bf16 path (in_dtype must be APB_BF16; K % 8 == 0 etc.). */
int apb_gemm_simt(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                  int trans_b, int epilogue, int in_dtype, int out_dtype, apb_stream_t stream);
int apb_gemm_tc(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                int trans_b, int epilogue, int in_dtype, int out_dtype, int split_k, apb_stream_t stream);
/* Same product for the forward-GEMM case (A [M,K] and B [N,K] both K-major, optional bias, bf16 or fp32 out) on CTA
 * PAIRS: tcgen05.mma.cta_group::2, 256 x 192 tile per pair, each CTA stages only half of the B tile (gemm_tc2.cu). */
int apb_gemm_tc_pair(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int out_dtype,
                     apb_stream_t stream);
/* split_k > 1 (wgrad: few output tiles, very long K): C receives split_k fp32 partial products [split_k][M][N] that
 * the caller sums in fixed order (apb_colsum over the split dim) -> deterministic; bias/epilogue must be 0/NULL. */
int apb_gemm_tc_suggest_split(int M, int N, int K);
/* out[i] = sum_s parts[s][i] in fixed order (n % 4 == 0): the reduction of the split-K partials. */
int apb_splitk_reduce(const float* parts, float* out, int splits, long long n, apb_stream_t stream);
/* apb_gemm_tc plus rowsum_parts[apb_gemm_tc_rowsum_slots(N, split_k)][M] (fp32): partial sums whose total over the slot
 * dim is sum_k A(m,k).  For the wgrad GEMM of nn.Linear (A = dY^T, models/volo.py:67-71 etc. backward) that is the bias
 * gradient: it is accumulated on the tensor pipe (one extra 128x16x16 MMA per k-step against a tile of ones, shared
 * between the n-tiles of an m-tile) from the dY tiles the GEMM has already staged in shared memory, so no separate pass
 * over dY is needed.  rowsum_parts may be NULL (== apb_gemm_tc). */
int apb_gemm_tc_rowsum(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                       int trans_b, int epilogue, int in_dtype, int out_dtype, int split_k, float* rowsum_parts,
                       apb_stream_t stream);
int apb_gemm_tc_rowsum_slots(int N, int split_k);
/* apb_splitk_reduce over two partial sets in one launch (weight-gradient tiles + the bias-gradient row sums);
 * n2 == 0 disables the second job. */
int apb_splitk_reduce2(const float* parts, float* out, long long n, int splits, const float* parts2, float* out2, long long n2,
                       int splits2, apb_stream_t stream);

/* ---- multi-head self-attention core  softmax(q k^T * scale) v  (models/volo.py:188-197)
 * qkv [B,N,3*heads*D] laid out (3, heads, D) per token; out [B,N,heads*D]; lse [B,heads,N] fp32 (saved for bwd).
 * bwd: dqkv same layout as qkv; workspace: B*heads*N floats (row dots D_i). */
int apb_mhsa_fwd(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, int dtype,
                 apb_stream_t stream);
/* tcgen05 / TMEM / TMA kernels for head_dim 32 and N <= 224 (attention_tc.cu): S (and dP) accumulate in TMEM, one TMEM
 * lane per query row (plain two-pass row softmax, no shuffles), P fed back as the TMEM A operand (forward) or through
 * 64B-swizzled shared tiles read K-major and MN-major (backward: dV, dK, dQ in ONE kernel after a row-dot pre-pass;
 * workspace: B*heads*N floats).
 * Return APB_ERR_UNSUPPORTED outside that envelope.  apb_mhsa_fwd / apb_mhsa_bwd use them by default for APB_BF16. */
int apb_mhsa_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, apb_stream_t stream);
int apb_mhsa_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace, int B,
                    int N, int heads, int D, float scale, apb_stream_t stream);
/* flash-style mma.sync kernels (attention_mma.cu; bf16, head_dim 32 or 64, any N that fits shared memory): scores in
 * registers, online softmax; bwd = row-dot + dQ + dK/dV kernels.  workspace: B*heads*N floats. */
int apb_mhsa_fwd_mma(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, apb_stream_t stream);
int apb_mhsa_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                     int B, int N, int heads, int D, float scale, apb_stream_t stream);
/* `_simt`: CUDA-core fp32-exact kernels (parity mode, any D <= 64).  The un-suffixed entries pick, for APB_BF16, the
 * tcgen05 kernels above (D == 32, N <= 224) or the flash-style mma.sync kernels (D == 64, or longer sequences), and the
 * SIMT ones for APB_F32. */
int apb_mhsa_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, int dtype,
                      apb_stream_t stream);
int apb_mhsa_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* workspace, int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream);
int apb_mhsa_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                 int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream);

/* ---- class attention core (cls query vs all tokens)  (models/volo.py:264-275)
 * q [B,heads*D]; kv [B,N,2*heads*D] laid out (2, heads, D); out [B,heads*D]. */
int apb_class_attn_fwd(const void* q, const void* kv, void* out, int B, int N, int heads, int D, float scale, int dtype,
                       apb_stream_t stream);
int apb_class_attn_bwd(const void* q, const void* kv, const void* dout, void* dq, void* dkv, int B, int N, int heads,
                       int D, float scale, int dtype, apb_stream_t stream);
/* the same with the keys in two buffers -- kv_cls [B, 2*heads*D] = the class token's own k / v row (key 0), kv_tok
 * [B, N-1, 2*heads*D] = the patch tokens (keys 1 .. N-1): the caller keeps [cls] and [tokens] apart instead of building
 * cat([cls, tokens]) for every ClassBlock (models/volo.py:300-308). */
int apb_class_attn_fwd_split(const void* q, const void* kv_cls, const void* kv_tok, void* out, int B, int N, int heads, int D,
                             float scale, int dtype, apb_stream_t stream);
int apb_class_attn_bwd_split(const void* q, const void* kv_cls, const void* kv_tok, const void* dout, void* dq, void* dkv_cls,
                             void* dkv_tok, int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream);

/* ---- elementwise / layout kernels
 * avgpool2: AvgPool2d(2,2,ceil_mode=True) on NHWC (models/volo.py:75,87) and its transpose.
 * flip_in_box: mix-token / un-mix (models/volo.py:655-658, 687-689): inside rows [r0,r1) x cols [c0,c1) of the
 *   [B,H,W,C] grid take sample B-1-b; self-inverse, so the backward is the same call on the gradient.
 * patchify: NHWC [B,H,W,C] -> rows [B*(H/p)*(W/p), p*p*C] with K order (kh,kw,c) (conv p x p stride p as a GEMM,
 *   models/volo.py:370-373, 389); unpatchify is the inverse permutation.
 * bicubic_resize: pos-embed resize (models/volo.py:580-596): src fp32 [h,w,C] -> dst fp32 [h0,w0,C],
 *   scale_factor=(h0+0.1)/h semantics, A=-0.75; bicubic_resize_bwd is its transpose (dst grad -> src grad).
 * add_bcast: out[b,i] = x[b,i] + p[i] (pos-embed broadcast add, in/out dtypes may differ).  cast: dtype conversion. */
int apb_avgpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, apb_stream_t stream);
int apb_avgpool2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int accumulate, int dtype, apb_stream_t stream);
int apb_flip_in_box(const void* x, void* y, int B, int H, int W, int C, int r0, int c0, int r1, int c1, int dtype,
                    apb_stream_t stream);
int apb_flip_in_box_dev(const void* x, void* y, int B, int H, int W, int C, const int* box_dev, int box_scale, int dtype,
                        apb_stream_t stream);   /* box (r0,c0,r1,c1) * box_scale read from device memory */
int apb_patchify(const void* x, void* rows, int B, int H, int W, int C, int p, int dtype, apb_stream_t stream);
int apb_unpatchify(const void* rows, void* x, int B, int H, int W, int C, int p, int dtype, apb_stream_t stream);
/* im2col of a small-channel convolution input (the 7x7 stride-2 stem conv of PatchEmbed, models/volo.py:352-353):
 * col [B*OH*OW, Kpad] bf16, column k = c*KH*KW + ky*KW + kx (== the row layout of the nn.Conv2d weight), zero-padded to
 * Kpad (multiple of 8) and outside the image.  x is read through element strides (sb, sc, sh, sw): NCHW or NHWC,
 * fp32 or bf16.  The conv itself is then one tcgen05 GEMM (apb_gemm_tc) and its weight gradient another. */
int apb_im2col(const void* x, void* col, int B, int C, int H, int W, int KH, int KW, int stride, int pad, int Kpad,
               long long sb, long long sc, long long sh, long long sw, int in_dtype, apb_stream_t stream);
/* Input resolution switch of the progressive schedule: F.interpolate(x, size=(OH, OW), mode='bilinear',
 * align_corners=False) (main_prog.py:973-974, 1910) on `planes` = B*C contiguous H x W fp32 planes (NCHW);
 * dst fp32 or bf16. */
int apb_bilinear_resize(const float* src, void* dst, long long planes, int H, int W, int OH, int OW, int out_dtype,
                        apb_stream_t stream);
int apb_bicubic_resize(const float* src, float* dst, int h, int w, int h0, int w0, int C, apb_stream_t stream);
int apb_bicubic_resize_bwd(const float* ddst, float* dsrc, int h, int w, int h0, int w0, int C, apb_stream_t stream);
int apb_add_bcast(const void* x, const float* p, void* out, long long batch, long long inner, int in_dtype,
                  int out_dtype, apb_stream_t stream);
int apb_cast(const void* in, void* out, long long n, int in_dtype, int out_dtype, apb_stream_t stream);
/* out[b,i] = in[b,i] * rs[b] (rs NULL = 1) with dtype conversion; residual_add: out = x + rs[b]*r. */
int apb_scale_cast(const void* in, const float* rs, void* out, long long batch, long long inner, int in_dtype,
                   int out_dtype, apb_stream_t stream);
int apb_residual_add(const void* x, const void* r, const float* rs, void* out, long long batch, long long inner,
                     int x_dtype, int r_dtype, int out_dtype, apb_stream_t stream);
int apb_add(const void* a, const void* b, void* out, long long n, int dtype, apb_stream_t stream);
int apb_gelu_fwd(const void* x, void* y, long long n, int dtype, apb_stream_t stream);
int apb_gelu_bwd(const void* x, const void* dy, void* dx, long long n, int dtype, apb_stream_t stream);

/* ---- train-mode BatchNorm2d + ReLU on channels-last activations (PatchEmbed stem, models/volo.py:355-368)
 * x,y,dy,dx [rows, C] NHWC-flattened (dtype); gamma,beta,mean,invstd,running_* fp32 [C]; C % 8 == 0.
 * fwd: use_batch_stats=1 computes batch mean / invstd (saved for backward) and updates the running statistics like
 * nn.BatchNorm2d (momentum, unbiased running variance); use_batch_stats=0 applies the given mean/invstd (eval).
 * bwd: g = dy * (y > 0); dbeta = sum g; dgamma = sum g*xhat; dx = gamma*invstd*(g - dbeta/rows - xhat*dgamma/rows).
 * workspace: apb_bn_workspace_floats(rows, C) floats. */
long long apb_bn_workspace_floats(long long rows, int C);
int apb_bn_relu_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* invstd,
                    float* running_mean, float* running_var, float momentum, float eps, int use_batch_stats,
                    float* workspace, long long rows, int C, int dtype, apb_stream_t stream);
/* backward: the ReLU mask is recomputed from x with the forward's expression when beta is given (y may then be NULL and is
 * not read: a third less traffic); with beta == NULL the mask is read from y. */
int apb_bn_relu_bwd(const void* x, const void* y, const void* dy, const float* gamma, const float* beta, const float* mean,
                    const float* invstd, void* dx, float* dgamma, float* dbeta, float* workspace, long long rows, int C,
                    int dtype, apb_stream_t stream);

/* ---- fused AdamW + k EMA updates + bf16 shadow weights in one pass over the parameters
 * (timm create_optimizer('adamw') + 4 x ModelEmaV2.update, main_prog.py:1019-1033)
 * p,g,m,v fp32 [n]; ema: array (device) of n_ema fp32 pointers, decay: host array of n_ema floats (<= 8);
 * shadow: optional bf16 copy of the updated parameters (NULL to skip). */
int apb_adamw_ema(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev, float beta1,
                  float beta2, float eps, float weight_decay, float* const* ema_ptrs_host, const float* decay_host,
                  int n_ema, void* shadow_bf16, apb_stream_t stream);
/* hyper_dev: DEVICE array {lr, 1-beta1^t, sqrt(1-beta2^t)} so that a captured CUDA graph replays with fresh values. */

/* number of kernel launches issued through this library since load (bench.py's gpu_launches evidence) */
long long apb_launch_count(void);
/* number of bf16 calls whose tensor-core kernel declined the shape (APB_ERR_UNSUPPORTED) and that were served by a
 * CUDA-core kernel instead; each one also prints a line on stderr.  bench.py asserts / reports it (must stay 0 on the
 * benchmark configurations). */
long long apb_fallback_count(void);
/* programmatic dependent launch: the hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization so
 * that their prologue overlaps the previous kernel's tail (each one executes griddepcontrol.wait before its first
 * global-memory access).  Off by default (saves 0.7-1.6 us per launch in chains of one kernel, neutral on the whole
 * training step); apb_set_pdl(1) turns it on. */
void apb_set_pdl(int on);
int apb_get_pdl(void);
/* diagnostic (tools/gemm_bound.py): dbg bit 0 = no TMA loads, bit 1 = no MMAs, bit 2 = no stores in the tcgen05 GEMM;
 * five_stage = 0 selects the four-stage 128x192 kernels for A/B timing.  Defaults: (0, 1). */
void apb_debug_gemm_switches(int dbg, int five_stage);

/* diagnostic (tools/umma_probe.py): D[128,32] = A[128,64] * B[64, off:off+32] through ONE descriptor convention
 * (mode 0..3, see csrc/umma_probe.cu); pins the shared-memory layouts the attention kernels rely on. */
int apb_debug_umma_probe(const void* A, const void* Bm, float* D, int mode, int off_elems, apb_stream_t stream);
/* diagnostic (tools/umma_timing.py): SM cycles for `reps` back-to-back tcgen05.mma of one operand-layout kind
 * (csrc/umma_timing.cu); out4 = {issue cycles, issue + completion cycles} x {warm-up, measured}. */
int apb_debug_umma_timing(long long* out4, int kind, int N, int reps, apb_stream_t stream);
/* diagnostic (tools/mhsa_trace.py): device buffer [18][1024] int64; CTA 0 of the following apb_mhsa_*_tc launches logs
 * (event id << 48 | SM clock) per warp into it.  NULL switches tracing off. */
int apb_debug_mhsa_trace(long long* buf);

#ifdef __cplusplus
}
#endif
#endif /* AUTOPROG_B200_H_ */
