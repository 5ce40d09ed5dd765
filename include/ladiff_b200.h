/* ladiff_b200 — C-ABI of the B200-native LaDiffCodec sampling path.
 *
 * The reference (haiciyang/LaDiffCodec) is pure PyTorch and has no FFI: its "operator API" is
 * nn.Module.__call__ on the objects srcs/sample.py builds.  Each entry point below
 * replaces one of those calls; the reference interface it stands in for is cited beside it
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stubs a reference
 * maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative LADIFF_ERR_* otherwise; ladiff_last_error()
 *    gives the message (thread-local).  No exceptions cross the boundary.
 *  - all tensor pointers are DEVICE pointers owned by the caller unless the parameter says "host";
 *    fp32, contiguous, NCL exactly as the reference passes them ([B,C,L]); indices are int64.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls only enqueue
 *    work; they do not synchronise unless stated.
 *  - `ws` is caller-owned scratch of at least ladiff_workspace_bytes() bytes, 256-byte aligned.
 *    The library allocates device memory only in ladiff_load_weight / ladiff_finalize.
 *  - one handle == one `DiffAudioRep` (srcs/model.py:32).  Handles are independent; a handle must
 *    not be used from two threads at once.
 */
#ifndef LADIFF_B200_H_
#define LADIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LADIFF_OK 0
#define LADIFF_ERR_ARG -1        /* bad argument / shape */
#define LADIFF_ERR_STATE -2      /* call order (e.g. not finalized) */
#define LADIFF_ERR_KEY -3        /* missing / unexpected / mis-shaped checkpoint key (strict load) */
#define LADIFF_ERR_CUDA -4       /* CUDA runtime / driver error */
#define LADIFF_ERR_UNSUPPORTED -5 /* constructor flag outside the sampling path */
#define LADIFF_ERR_WORKSPACE -6  /* workspace too small or misaligned */

#define LADIFF_MAX_RATIOS 8

/* Mirrors the keyword arguments of DiffAudioRep.__init__ (srcs/model.py:34) that shape the
 * sampling path.  Flags the path does not support (use_film, self_condition, qtz_condition,
 * unet_scale_x, run_vae, model_type != 'unet') are rejected in the Python host. */
typedef struct LadiffConfig {
  int32_t rep_dims;            /* 128 */
  int32_t diff_dims;           /* 256 */
  int32_t n_filters;           /* 32 */
  int32_t lstm_layers;         /* 2 (0 = no LSTM) */
  int32_t n_enc_ratios;        /* len(enc_ratios) */
  int32_t enc_ratios[LADIFF_MAX_RATIOS];
  int32_t quantization;        /* 1: has RVQ (the conditioning codec) */
  int32_t n_q;                 /* quantizers built (model.py:65) */
  int32_t n_q_used;            /* quantizers used by forward() at `bandwidth` (vq.py:86-98) */
  int32_t run_diff;            /* 1: has Unet1D + GaussianDiffusion1D */
  int32_t cond_channels;       /* 128 */
  int32_t n_upsampling_ratios; /* len(upsampling_ratios); 0 = None */
  int32_t upsampling_ratios[LADIFF_MAX_RATIOS];
  int32_t unet_scale_cond;     /* per-sample max-abs scaling of the upsampled cond (unet.py:417) */
  int32_t sample_rate;         /* 16000 */
  int32_t reserved[8];
} LadiffConfig;

typedef struct LadiffHandle LadiffHandle;

const char* ladiff_last_error(void);
/* ABI version of this header; bump on any signature change. */
int32_t ladiff_abi_version(void);
/* 16-bit type of the UNet's tensor-core operands and stored activations in this build: "f16" (default) or "bf16". */
const char* ladiff_act_dtype(void);

/* DiffAudioRep(**kwargs)                                              srcs/model.py:34-106 */
int32_t ladiff_create(const LadiffConfig* cfg, LadiffHandle** out);
int32_t ladiff_destroy(LadiffHandle* h);

/* load_model(model, path, strict=True) → model.load_state_dict        srcs/utils.py:98-108
 * One call per state-dict entry, name exactly as stored (after stripping 'module.').  `data`
 * may be a host or a device pointer (fp32, contiguous); the library copies it.  Keys that are
 * not part of the layout for this config fail with LADIFF_ERR_KEY (strict). */
int32_t ladiff_load_weight(LadiffHandle* h, const char* name, const float* data,
                           const int64_t* shape, int32_t ndim);
/* End of strict load: verifies every expected key arrived with the expected shape, then folds
 * (weight-norm g*v/||v||, weight standardisation, time-MLP → FiLM table, 16-bit K-major repack).
 * Synchronises the device. */
int32_t ladiff_finalize(LadiffHandle* h);
/* Number of keys the strict layout expects (745 for the README LaDiff model, 148 for the codec). */
int32_t ladiff_expected_keys(const LadiffHandle* h);
int32_t ladiff_expected_key_at(const LadiffHandle* h, int32_t i, const char** name,
                               int64_t* shape4, int32_t* ndim);

/* Scratch needed by any call below for batch B and waveform length T (samples). */
int64_t ladiff_workspace_bytes(const LadiffHandle* h, int32_t B, int32_t T);

/* model_for_cond.get_cond(wav)                                         srcs/model.py:223-231
 *   = SEANetEncoder (seanet.py:153) → ResidualVectorQuantizer.forward (vq.py:69-84).
 * wav [B,1,T] → cond [B,rep_dims,T/hop]; codes [n_q_used,B,T/hop] int64 (optional, may be NULL);
 * enc_out [B,rep_dims,T/hop] pre-quantisation encoder output (optional). */
int32_t ladiff_get_cond(LadiffHandle* h, const float* wav, int32_t B, int32_t T, float* cond,
                        int64_t* codes, float* enc_out, void* ws, int64_t ws_bytes, void* stream);

/* model.encoder(x)                                                     srcs/modules/seanet.py:153 */
int32_t ladiff_encode(LadiffHandle* h, const float* wav, int32_t B, int32_t T, float* z,
                      void* ws, int64_t ws_bytes, void* stream);

/* quantizer.decode(codes)  (codes-in wire format)  srcs/quantization/vq.py:109-113, core_vq.py:356-362
 * codes [n_q,B,F] int64 → quantized [B,rep_dims,F]. */
int32_t ladiff_rvq_decode(LadiffHandle* h, const int64_t* codes, int32_t n_q, int32_t B, int32_t F,
                          float* quantized, void* stream);
/* quantizer.encode(z) / forward(z): z [B,D,F] → codes [n_q,B,F], quantized [B,D,F] (either may be NULL) */
int32_t ladiff_rvq_encode(LadiffHandle* h, const float* z, int32_t n_q, int32_t B, int32_t F,
                          int64_t* codes, float* quantized, void* stream);

/* model.diff_model.upsampling_layers[i](x)                             srcs/modules/unet.py:372-377,
 *   SConvTranspose1d(causal=False)                                     srcs/modules/conv.py:252-274
 * x [B,C,Lin] → y [B,C,Lin*ratio_i]. */
int32_t ladiff_upsample_layer(LadiffHandle* h, int32_t i, const float* x, int32_t B, int32_t Lin,
                              float* y, void* stream);

/* model.diff_model(x, time, x_cond)  = Unet1D.forward                  srcs/modules/unet.py:422-469
 * x [B,rep_dims,L], time [B] int64, cond [B,cond_channels,F] (un-upsampled; process_cond is applied
 * inside exactly like the reference) → eps [B,rep_dims,L]. */
int32_t ladiff_unet_forward(LadiffHandle* h, const float* x, const int64_t* time, const float* cond,
                            int32_t B, int32_t L, int32_t F, float* eps, void* ws, int64_t ws_bytes,
                            void* stream);

/* model.diffusion.halfway_sampling(img, t, condition) / p_sample loop  srcs/losses/ddpm_loss.py:244-251,370-385
 * Runs DDPM steps i = t_start-1 … t_start-n_steps on x [B,rep_dims,L] in place, conditioned on
 * cond [B,cond_channels,F].  noise: pre-drawn [n_noise,B,rep_dims,L] consumed in loop order (one
 * per step with i > 0), or NULL → Philox draws from `seed` (documented to differ from torch's stream).
 * process_cond is hoisted out of the loop (it depends only on cond). */
int32_t ladiff_ddpm_steps(LadiffHandle* h, float* x, const float* cond, const float* noise,
                          int64_t n_noise, uint64_t seed, int32_t t_start, int32_t n_steps,
                          int32_t B, int32_t L, int32_t F, void* ws, int64_t ws_bytes, void* stream);

/* In-kernel noise (noise == NULL in the sampler entry points) is Philox4x32-10 keyed by (seed, absolute timestep, GLOBAL clip
 * index, element): `clip_offset` is the global index of this handle's clip 0 (a rank's first clip when a job is sharded), so a
 * job draws the same noise however it is split over calls, streams or GPUs.  Default 0. */
int32_t ladiff_set_clip_offset(LadiffHandle* h, uint64_t clip_offset);

/* model.diffusion.ddim_sample's loop                                   srcs/losses/ddpm_loss.py:268-303
 * times: HOST array of n_pairs+1 decreasing timesteps (the reference's reversed linspace(-1, T-1, S+1).int(); the last may be
 * -1 = "return x_start").  For every pair (time, time_next): eps = Unet1D(x, time, cond); x0 = clamp(a x - b eps, -1, 1);
 * time_next < 0 → x = x0, else x = x0 sqrt(ac[next]) + c eps + sigma z with sigma = eta sqrt((1-ac/ac_next)(1-ac_next)/(1-ac)),
 * c = sqrt(1 - ac_next - sigma^2).  noise: pre-drawn [n_noise,B,rep_dims,L], one per pair with time_next >= 0 (the reference
 * draws it even when eta == 0), or NULL → in-kernel generator.  x [B,rep_dims,L] in place. */
int32_t ladiff_ddim_steps(LadiffHandle* h, float* x, const float* cond, const int32_t* times, int32_t n_pairs, double eta,
                          const float* noise, int64_t n_noise, uint64_t seed, int32_t B, int32_t L, int32_t F, void* ws,
                          int64_t ws_bytes, void* stream);
/* times_out[0..sampling_timesteps] = the reference's DDIM time grid (ddpm_loss.py:273-274), host array. */
int32_t ladiff_ddim_times(int32_t total_timesteps, int32_t sampling_timesteps, int32_t* times_out);

/* torch.randn(shape) / torch.rand(shape) stand-ins for the samplers' initial draws (ddpm_loss.py:256,277,336) when the caller
 * does not pass its own: x [B, n_per_clip] ← N(0,1) (uniform = 0) or U[0,1) (uniform = 1) from the in-kernel generator. */
int32_t ladiff_randn(LadiffHandle* h, float* x, int32_t B, int64_t n_per_clip, uint64_t seed, int32_t uniform, void* stream);

/* model.diffusion.q_sample(x_start, t, noise)                          srcs/losses/ddpm_loss.py:387-393
 * out = sqrt_alphas_cumprod[t_b] x_start + sqrt_one_minus_alphas_cumprod[t_b] noise; t [B] int64. ws: >= 1 KB + 4 B bytes. */
int32_t ladiff_q_sample(LadiffHandle* h, const float* x_start, const int64_t* t, const float* noise, float* out, int32_t B,
                        int64_t n_per_clip, void* ws, int64_t ws_bytes, void* stream);

/* x ← a x + b y (y NULL: x ← a x).  infilling's `(1 - lam) * img + lam * infill_img` (ddpm_loss.py:360,364) and
 * DiffAudioRep.scaling's division by the global constant (model.py:137-139).  n elements, fp32, in place. */
int32_t ladiff_axpby(float* x, double a, const float* y, double b, int64_t n, void* stream);

/* model.diffusion(x_start, cond, t, noise) = GaussianDiffusion1D.forward → p_losses   srcs/losses/ddpm_loss.py:404-450
 * (loss_type l1, objective pred_noise; forward only — the reference's two UNet evaluations are the same tensor).
 * x_t = q_sample(x_start, t, noise); model_out = Unet1D(x_t, t, cond); predicted_x_start = a_t x_t - b_t model_out (no clamp);
 * loss [1] = mean_b( mean|model_out - noise| * p2_loss_weight[t_b] ).  pred_x_start, model_out optional [B,rep_dims,L]. */
int32_t ladiff_p_losses(LadiffHandle* h, const float* x_start, const int64_t* t, const float* cond, const float* noise, int32_t B,
                        int32_t L, int32_t F, float* loss, float* pred_x_start, float* x_t, float* model_out, void* ws,
                        int64_t ws_bytes, void* stream);

/* sdr_loss(est, target) = ClippedSDR(MultiSrcNegSDR("sdsdr"))          srcs/losses/losses_fn.py:54-66 (asteroid's sd-sdr)
 * est, target [B, n] (one source per clip) → out [B] = max(-sdsdr, clip_value). */
int32_t ladiff_sdsdr(const float* est, const float* target, float* out, int32_t B, int64_t n, double clip_value, void* stream);

/* model.decoder(z)  = SEANetDecoder.forward                            srcs/modules/seanet.py:246-248
 * z [B,rep_dims,L] → wav [B,1,L*hop]. */
int32_t ladiff_decode(LadiffHandle* h, const float* z, int32_t B, int32_t L, float* wav,
                      void* ws, int64_t ws_bytes, void* stream);

/* The per-file body of synthesis()                                     srcs/sample.py:94-134
 * get_cond(cond model) → upsampling_layers → img /= max|img| → halfway_sampling(t = n_steps) →
 * decoder → /= std → /= max, with every whole-tensor normalisation applied per clip (the reference
 * runs B = 1).  wav_in [B,1,T] → wav_out [B,1,T].  latent_out [B,rep_dims,L] optional. */
int32_t ladiff_synthesize(LadiffHandle* model, LadiffHandle* cond_model, const float* wav_in,
                          int32_t B, int32_t T, int32_t n_steps, const float* noise, int64_t n_noise,
                          uint64_t seed, float* wav_out, float* latent_out, void* ws, int64_t ws_bytes,
                          void* stream);

/* The per-file body with the reference's DDIM sampler (srcs/losses/ddpm_loss.py:268-303) in place of halfway_sampling:
 * get_cond → ddim_sample((B, rep_dims, L), condition) with `sampling_timesteps` steps from N(0, I) → decoder → normalise.
 * init_noise [B,rep_dims,L] (NULL: in-kernel generator); noise as in ladiff_ddim_steps. */
int32_t ladiff_synthesize_ddim(LadiffHandle* model, LadiffHandle* cond_model, const float* wav_in, int32_t B, int32_t T,
                               int32_t sampling_timesteps, double eta, const float* init_noise, const float* noise, int64_t n_noise,
                               uint64_t seed, float* wav_out, float* latent_out, void* ws, int64_t ws_bytes, void* stream);

/* The same, starting from the codec payload (the receiver of the 1.5/3 kbps stream): codes [n_q,B,F] int64 RVQ
 * indices (quantizer.decode, srcs/quantization/vq.py:108-113 → core_vq.py:356-362) instead of a waveform;
 * decodes T = 320·F samples.  Bit-identical to ladiff_synthesize on the waveform those codes came from. */
int32_t ladiff_synthesize_codes(LadiffHandle* model, LadiffHandle* cond_model, const int64_t* codes, int32_t n_q,
                                int32_t B, int32_t F, int32_t n_steps, const float* noise, int64_t n_noise,
                                uint64_t seed, float* wav_out, float* latent_out, void* ws, int64_t ws_bytes,
                                void* stream);

/* Scratch needed by ladiff_synthesize / ladiff_synthesize_codes for this pair of models. */
int64_t ladiff_synthesize_workspace_bytes(const LadiffHandle* model, const LadiffHandle* cond_model,
                                          int32_t B, int32_t T);

/* Per-clip script-level normalisations                                  srcs/sample.py:129,133-134
 * mode 0: x /= max|x| + 1e-8 ; mode 1: x /= std_unbiased(x) + 1e-8 then x /= max|x| + 1e-8 ;
 * mode 2: x /= max|x| + 1e-20 (Unet1D.scaling, srcs/modules/unet.py:401-403).  x [B,n] in place. */
int32_t ladiff_normalize_clips(float* x, int32_t B, int64_t n, int32_t mode, void* stream);

/* ---- operator-level entry points (tests, profiling) ------------------------------------------ */
/* Channels-last 16-bit (ladiff_act_dtype()) Conv1d on the tcgen05 path: x [B,L,Cin] f16, w [Cout,Cin,k] fp32 (PyTorch
 * layout), zero padding (k-1)/2, stride 1 → y [B,L,Cout] (f16, or fp32 if y_f32).  impl 0 = tcgen05,
 * 1 = SIMT check kernel (same packed operands), 2 = tcgen05 with per-tap activation tiles, 3 = tcgen05 positions-on-M
 * kernel, 4 = the same with a cta_group::2 CTA pair per tile,
 * 5 = channels-on-M kernel shaped for two CTAs per SM (small-K convs only), 6 = channels-on-M kernel as CTA pairs along N that
 * share every weight tile through TMA multicast.  Allocates and frees its own scratch; synchronises. */
int32_t ladiff_op_conv1d_cl(const void* x_h16, const float* w, const float* bias, int32_t B, int32_t L,
                            int32_t Cin, int32_t Cout, int32_t k, void* y, int32_t y_f32, int32_t impl,
                            float* gn_stats /* [B,Cout/32,2] or NULL */);
/* Attention core of the mid block (srcs/modules/unet.py:238-245) on channels-last 16-bit qkv [B,L,384] (q | k | v, 4 heads x 32) →
 * out [B,L,128].  impl 0 = the UNet's choice, 1 = tiled SIMT, 2 = SIMT with keys in shared memory (L <= 512), 3 = tcgen05 (QK^T, PV). */
int32_t ladiff_op_fullattn(const void* qkv_h16, void* out_h16, int32_t B, int32_t L, int32_t impl);
/* Operator-level timing of the same conv (kernel tuning, profiles/conv_sweep.py): one plan, `warm` untimed + `iters` timed launches
 * between two CUDA events; ms_out[0] = mean ms per launch; want_* override the tile shape (0 = cost model); label receives the plan. */
int32_t ladiff_op_conv1d_bench(const void* x_h16, const float* w, const float* bias, int32_t B, int32_t L, int32_t Cin, int32_t Cout,
                               int32_t k, int32_t want_nt, int32_t want_nclip, int32_t want_two, int32_t want_t, int32_t stats_on,
                               int32_t warm, int32_t iters, float* ms_out, char* label, int32_t label_cap);
/* Selects the conv implementation used inside the UNet: 0 = tcgen05 (default), 1 = SIMT check kernel. */
int32_t ladiff_set_conv_impl(LadiffHandle* h, int32_t impl);
/* Timing ablation for bench.py's roofline: UNet evaluations of this handle skip the kernel classes in `mask` (1 GroupNorm-apply,
 * 2 LayerNorm, 4 attention cores, 8 1x1 convs, 16 all other convs) — results are garbage while mask != 0.  The difference between
 * the replay time with and without a class is that class's time inside the CUDA-graph replay (PDL overlap included). */
int32_t ladiff_set_skip_ops(LadiffHandle* h, int32_t mask);
/* Profiling for bench.py's roofline: when on, CUDA events are recorded on the launching stream around every
 * tcgen05 conv launch of a UNet evaluation.  ladiff_profile_report (synchronises) returns for the most recent
 * evaluation: out4 = {conv ms, conv algorithmic FLOPs, conv launches, whole-evaluation ms}. */
int32_t ladiff_set_profiling(LadiffHandle* h, int32_t on);
int32_t ladiff_profile_report(LadiffHandle* h, double* out4);
/* Text table of the same evaluation, one line per conv launch: "<ms> <algorithmic GFLOP> <shape label>". */
int32_t ladiff_profile_dump(LadiffHandle* h, char* buf, int64_t cap);
/* Kernel launches issued by this handle since the last call (for bench.py's gpu_launches). */
int64_t ladiff_take_launch_count(LadiffHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* LADIFF_B200_H_ */
