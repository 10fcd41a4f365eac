/*
 * beatthis.h -- C ABI of libbeatthis_sm100.so, the B200 (sm_100a) implementation of the
 * CPJKU/beat_this Audio -> Frames -> Beats inference path.
 *
 * The reference has no FFI: its boundary is the Python API of beat_this/inference.py
 * (Spect2Frames / Audio2Frames / Audio2Beats / File2Beats).  Each entry point below names
 * the reference function (file:line under the reference tree) whose arithmetic it
 * replaces.  beat_this_b200/_lib.py binds these with ctypes; INTEGRATION.md shows the stub
 * a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; no torch / CUDA types in signatures (streams travel as void*).
 *  - `*_dev` pointers are device pointers on the context's GPU, `*_host` are host pointers.
 *  - ragged batches are CSR style: `offsets[n+1]` (host, int64) into a concatenated buffer.
 *  - all work is enqueued on the given CUDA stream (cudaStream_t as void*, NULL = default
 *    stream); no hidden device synchronisation except where stated.
 *  - return value: 0 = ok, negative = error (bt_last_error() gives the text).  Nothing
 *    throws across the ABI.  There is NO CPU fallback: without a CUDA device every compute
 *    entry point fails with BT_ERR_CUDA.
 *  - one bt_ctx per GPU; a ctx is not thread-safe (one host thread per ctx).
 */
#ifndef BEATTHIS_H_
#define BEATTHIS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_OK 0
#define BT_ERR_ARG (-1)
#define BT_ERR_CUDA (-2)
#define BT_ERR_STATE (-3)
#define BT_ERR_PARAM (-4)
#define BT_ERR_IO (-5)     /* file could not be opened / read */
#define BT_ERR_FORMAT (-6) /* not a RIFF/WAVE file this library decodes (caller falls back to another decoder) */

#define BT_DTYPE_F32 0  /* fp32 CUDA-core kernels: the reference's float16=False numerics   */
#define BT_DTYPE_H16 1 /* 16-bit tcgen05 tensor-core kernels, fp32 accumulate + fp32 residual stream.  Operand type
                        * fp16 (what the reference's float16=True autocasts to on CUDA, inference.py:245-246);
                        * bt_act_dtype() names it ("f16", or "bf16" for a -DBT_ACT_BF16 build) */

#define BT_SAMPLE_RATE 22050
#define BT_N_FFT 1024
#define BT_HOP 441
#define BT_N_MELS 128
#define BT_CHUNK 1500
#define BT_BORDER 6
#define BT_FPS 50

typedef struct bt_ctx bt_ctx;

/* BeatThis constructor arguments (reference beat_this/model/beat_tracker.py:39-49), as
 * filtered from the checkpoint's hyper_parameters by load_model (inference.py:72-78). */
typedef struct bt_hparams {
  int32_t spect_dim;            /* 128 */
  int32_t transformer_dim;      /* 512 (final*) / 128 (small*) */
  int32_t ff_mult;              /* 4 */
  int32_t n_layers;             /* 6 */
  int32_t head_dim;             /* 32 */
  int32_t stem_dim;             /* 32 */
  int32_t sum_head;             /* 1 */
  int32_t partial_transformers; /* 1 */
} bt_hparams;

/* Library / ABI version (major*100 + minor). */
int bt_version(void);

/* Operand type of the BT_DTYPE_H16 path in this build: "f16" or "bf16". */
const char* bt_act_dtype(void);

/* ---- model lifetime: replaces load_model (inference.py:56-87) ------------------------ */

/* Create a context on CUDA device `device_ordinal` for a model of shape `hp`.
 * compute_dtype: BT_DTYPE_F32 or BT_DTYPE_H16. */
int bt_create(bt_ctx** out, int device_ordinal, const bt_hparams* hp, int compute_dtype);

/* Upload one packed parameter (fp32, host memory, `count` elements) under `name`.
 * The host side (beat_this_b200/weights.py) folds eval-mode BatchNorm and the RMSNorm
 * gamma*sqrt(dim) factors into neighbouring weights and lays convolution / frontend.linear
 * weights out as GEMM operands; names and shapes are listed in DESIGN.md "Packed
 * parameters".  Replaces model.load_state_dict (inference.py:84). */
int bt_set_param(bt_ctx* ctx, const char* name, const float* data_host, int64_t count);

/* Check that every parameter the model shape needs is present, build 16-bit operand copies and TMA
 * descriptors.  After this the weights are immutable.  (model.to(device).eval(), :87) */
int bt_finalize(bt_ctx* ctx);

/* Free everything. */
void bt_destroy(bt_ctx* ctx);

/* Text of the last error on this ctx (or the last bt_create failure when ctx == NULL). */
const char* bt_last_error(const bt_ctx* ctx);

/* ---- host-side planning helpers (pure host, no GPU needed) ---------------------------------- */

/* Number of log-mel frames for `n_samples` samples: 1 + n/441 (torch.stft center=True,
 * preprocessing.py:43-59). */
int64_t bt_num_frames(int64_t n_samples);

/* split_piece (inference.py:100-135) for a piece of T frames with chunk 1500 / border 6 /
 * avoid_short_end: writes up to `cap` chunk starts and lengths, returns the chunk count
 * (or the count needed when cap is too small). */
int64_t bt_plan_chunks(int64_t T, int64_t* starts, int64_t* lens, int64_t cap);

/* ---- host front door: Audio2Frames.signal2spect's host half (inference.py:269-276) ------------------ */

#define BT_SIG_F32 0 /* float32 samples                                                         */
#define BT_SIG_F64 1 /* float64 samples (what the reference's load_audio returns)                */
#define BT_SIG_I16 2 /* int16 PCM, scaled by 1/32768 (== soundfile / torchaudio's float decode)  */

/* Mono mix + fp32 cast of n_clips host signals into one host buffer (normally pinned; it is the source of the single
 * H2D copy of a batch): clip i is signals[i], `frames[i]` x `channels[i]` samples, channels-last and contiguous, of
 * type dtypes[i]; its mono fp32 samples land at dst + dst_offsets[i].  The arithmetic is the reference's: the mean
 * over channels in the array's own floating type (numpy `signal.mean(1)`, inference.py:270-271; float64 for PCM),
 * then the cast of `torch.tensor(signal, dtype=torch.float32)` (:276).  Runs on n_threads host threads (<= 0: all);
 * no CUDA, no ctx. */
int bt_stage_audio(const void* const* signals, const int32_t* dtypes, const int64_t* frames,
                   const int32_t* channels, int32_t n_clips, float* dst, const int64_t* dst_offsets,
                   int32_t n_threads);

/* RIFF/WAVE front door of load_audio (preprocessing.py:6-24) for the batched File2Beats path. */
typedef struct bt_wav_info {
  int32_t sample_rate;
  int32_t channels;
  int32_t bytes_per_sample; /* 1, 2, 3, 4 (PCM) or 4, 8 (IEEE float) */
  int32_t is_float;
  int64_t frames;
  int64_t data_offset;      /* byte offset of the first sample in the file */
} bt_wav_info;

/* Parse the header of a WAV file (PCM 8/16/24/32-bit, IEEE float 32/64-bit, incl. WAVE_FORMAT_EXTENSIBLE).
 * BT_ERR_IO: cannot open; BT_ERR_FORMAT: not such a file. */
int bt_wav_probe(const char* path, bt_wav_info* info);

/* Decode n_files probed WAV files on n_threads host threads: samples -> float64 as soundfile would return them
 * (PCM / 2^(bits-1)), mean over channels in float64, fp32 cast -- i.e. load_audio + the host half of
 * signal2spect -- written to dst + dst_offsets[i] (infos[i].frames samples each).  status[i] (optional) receives
 * BT_OK / BT_ERR_IO per file; files that fail are zero-filled and the call returns BT_ERR_IO. */
int bt_stage_wav_files(const char* const* paths, const bt_wav_info* infos, int32_t n_files, float* dst,
                       const int64_t* dst_offsets, int32_t n_threads, int32_t* status);

/* ---- the hot path ------------------------------------------------------------------------- */

/* LogMelSpect.forward (preprocessing.py:56-59) for n_clips mono 22.05 kHz clips.
 * audio_dev: concatenated fp32 samples; sample_offsets_host[n_clips+1].
 * spect_dev: out, concatenated [T_i,128] fp32 with T_i = bt_num_frames(len_i), laid out
 * at frame_offsets_host[i] (frames; frame_offsets_host[n_clips+1]). */
int bt_logmel(bt_ctx* ctx, const float* audio_dev, const int64_t* sample_offsets_host,
              int32_t n_clips, float* spect_dev, const int64_t* frame_offsets_host,
              void* stream);

/* Resample front door of Audio2Frames.signal2spect (inference.py:274-275:
 * `soxr.resample(signal, in_rate=sr, out_rate=22050)`), as a device polyphase FIR:
 *   out[n] = sum_k coef[(n*M) mod L][k] * in[floor(n*M/L) - K/2 + 1 + k]   (zeros outside a clip)
 * with sr_out/sr_in = L/M in lowest terms.  coef_dev: [L][K] fp32 bank (the host side designs it,
 * beat_this_b200/preprocessing.py: Kaiser-windowed sinc to soxr-HQ-like targets; parity with
 * soxr itself is unpinned).  in/out: concatenated fp32 clips with host offset arrays
 * [n_clips+1]; out lengths are the caller's (normally round(len*L/M)). */
int bt_resample(bt_ctx* ctx, const float* audio_in_dev, const int64_t* in_offsets_host,
                int32_t n_clips, const float* coef_dev, int32_t L, int32_t M, int32_t K,
                float* audio_out_dev, const int64_t* out_offsets_host, void* stream);

/* Host-side Viterbi of the bar-pointer HMM behind Postprocessor(type="dbn") (model/postprocessor.py:29-37,170:
 * madmom DBNDownBeatTrackingProcessor; restated in beat_this_b200/dbn.py, parity with madmom unpinned).
 * No CUDA, no ctx: thread-safe.  log_dens [T][3] = log densities of (no beat, beat, downbeat); the state space is
 * `beats` beats x the positions of every tempo `intervals[n_int]` (frames per beat); log_tempo [n_int][n_int] =
 * log P(tempo f -> tempo k) at a beat boundary; pointers [S] = density column of every state.
 * path_out [T] receives the most probable state sequence, *logp_out its log-probability. */
int bt_dbn_viterbi(const double* log_dens, int64_t T, int32_t beats, int32_t n_int,
                   const int32_t* intervals, const double* log_tempo, const int32_t* pointers,
                   int64_t* path_out, double* logp_out);

/* The whole DBN post-processing step of Postprocessor.postp_dbn (model/postprocessor.py:138-173) for many pieces
 * at once, multi-threaded on the host: activations [total_frames][2] = (beat-but-not-downbeat, downbeat)
 * probabilities as the reference builds them (:159-167), pieces at frame_offsets[n_clips+1].  Model parameters as
 * in madmom's DBNDownBeatTrackingProcessor (reference values: beats_per_bar {3,4}, 55..215 BPM, num_tempi 60,
 * transition_lambda 100, observation_lambda 16, threshold 0.05, correct 1, fps 50).  Piece i writes
 * counts_out[i] (time [s], beat number) pairs at times_out / numbers_out + frame_offsets[i]; downbeats are the
 * entries with number 1.  n_threads <= 0: hardware concurrency. */
int bt_dbn_track(const double* activations, const int64_t* frame_offsets, int32_t n_clips,
                 const int32_t* beats_per_bar, int32_t n_bar_lengths, double min_bpm, double max_bpm,
                 int32_t num_tempi, double transition_lambda, double observation_lambda,
                 double threshold, int32_t correct, double fps, int32_t n_threads,
                 double* times_out, int32_t* numbers_out, int64_t* counts_out);

/* Spect2Frames.spect2frames (inference.py:244-254): split_piece -> BeatThis.forward on
 * every chunk -> aggregate_prediction(keep_first).  spect_dev as produced by bt_logmel.
 * beat_dev / downbeat_dev: out, fp32 logits, concatenated with the same frame offsets. */
int bt_spect2frames(bt_ctx* ctx, const float* spect_dev, const int64_t* frame_offsets_host,
                    int32_t n_clips, float* beat_dev, float* downbeat_dev, void* stream);

/* BeatThis.forward (model/beat_tracker.py:188-192) on n_chunks spectrogram chunks of chunk_frames (<= 1500) frames
 * each, chunks_dev = [n_chunks, chunk_frames, 128] fp32: no chunk planning, no borders cut -- the model call inside
 * split_predict_aggregate (inference.py:215), batched.  beat_dev / downbeat_dev: [n_chunks, chunk_frames] fp32. */
int bt_forward_chunks(bt_ctx* ctx, const float* chunks_dev, int32_t n_chunks, int32_t chunk_frames,
                      float* beat_dev, float* downbeat_dev, void* stream);

/* Audio2Frames.__call__ (inference.py:279-281) for already mono, 22.05 kHz fp32 audio:
 * bt_logmel + bt_spect2frames with the spectrogram kept in the ctx workspace. */
int bt_audio2frames(bt_ctx* ctx, const float* audio_dev, const int64_t* sample_offsets_host,
                    int32_t n_clips, float* beat_dev, float* downbeat_dev,
                    const int64_t* frame_offsets_host, void* stream);

/* Postprocessor("minimal") (model/postprocessor.py:85-136,176-197) on device.
 * Per clip i: beat_times_dev[i*max_peaks ..] (float64 seconds), n_beats_dev[i], same for
 * downbeats.  A clip with more than max_peaks peaks reports the true count (> max_peaks)
 * and stores the first max_peaks. */
int bt_peakpick(bt_ctx* ctx, const float* beat_dev, const float* downbeat_dev,
                const int64_t* frame_offsets_host, int32_t n_clips, double* beat_times_dev,
                int32_t* n_beats_dev, double* down_times_dev, int32_t* n_down_dev,
                int32_t max_peaks, void* stream);

/* ---- introspection / tuning ----------------------------------------------------------------- */

/* Upper bound on the chunks processed per wave (default 128; one wave = one launch of every
 * kernel of the forward pass).  The workspace (~46 MB per 1500-frame chunk on the 16-bit path) grows on
 * demand up to this many chunks. */
int bt_set_wave_chunks(bt_ctx* ctx, int32_t chunks);

/* Number of kernel launches issued by this ctx since creation (bench.py "gpu_launches"). */
int64_t bt_launch_count(const bt_ctx* ctx);

/* Per-kernel-class device timing for bench.py's roofline line.  While enabled, one CUDA event
 * is recorded on the launch stream after every kernel launch; a launch's duration is the gap
 * to the previous event.  bt_profile_collect synchronises the device and folds the recorded
 * events into per-class totals; bt_profile_get(index) reads class `index` (name, total
 * milliseconds, launches), bt_profile_count the number of classes seen so far. */
int bt_profile_enable(bt_ctx* ctx, int enable);
int bt_profile_collect(bt_ctx* ctx);
int bt_profile_reset(bt_ctx* ctx);
int bt_profile_count(const bt_ctx* ctx);
int bt_profile_get(const bt_ctx* ctx, int index, char* name, int name_cap, double* total_ms,
                   int64_t* launches);

/* Test hook: after the next bt_spect2frames call on a single wave, copy the activation
 * named `tap` (see DESIGN.md "Taps") as fp32 into out_dev (capacity `cap` floats).
 * Returns the element count through *count.  Used only by tests/. */
int bt_debug_request_tap(bt_ctx* ctx, const char* tap, float* out_dev, int64_t cap);
int64_t bt_debug_tap_count(const bt_ctx* ctx);

/* Test hook: D[M,N] = A[M,K] * W[N,K]^T through the ctx's GEMM for its compute dtype
 * (fp32 inputs on device; the 16-bit path rounds A and W to its operand type first). */
int bt_debug_gemm(bt_ctx* ctx, const float* a_dev, const float* w_dev, float* d_dev,
                  int32_t M, int32_t N, int32_t K, void* stream);

/* Test hook: softmax(Q K^T / sqrt(32)) V for `seqs` sequences of length L and `heads`
 * heads of dim 32 through the ctx's time-direction attention kernel.  q/k/v/o_dev are
 * [seqs, L, heads*32] fp32. */
int bt_debug_attention(bt_ctx* ctx, const float* q_dev, const float* k_dev, const float* v_dev,
                       float* o_dev, int32_t seqs, int32_t L, int32_t heads, void* stream);

/* Profiling hook (tools/attn_ubench.py): time `iters` launches of the 16-bit time-direction attention kernel on
 * synthetic q|k|v of [seqs, L, heads*32]; variant < 0 keeps the default kernel, otherwise the template parameter V
 * of attn_tc48_kernel (kernels_attn.cu lists the compiled ones).  *ms_per_launch from CUDA events. */
int bt_debug_attention_time(bt_ctx* ctx, int32_t seqs, int32_t L, int32_t heads, int32_t variant,
                            int32_t iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* BEATTHIS_H_ */
