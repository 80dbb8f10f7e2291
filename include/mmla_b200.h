/*
 * mmla_b200.h — C-ABI of libmmla_b200.so: the B200 (sm_100a) implementation of the
 * mmla-audio analytics hot path.
 *
 * The reference (lizaibeim/mmla-audio) has no FFI: its boundary is the set of Python call
 * signatures its entry scripts use.  Each entry point below names the reference
 * call it replaces (paths relative to the reference root); the Python layer in
 * mmla_audio_b200/ binds these with ctypes and re-exposes the reference's own signatures
 * (INTEGRATION.md shows the stub a maintainer would add).
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch / C++ types.
 *   - All data pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *     enqueue work; they never synchronise the device.
 *   - Return value: 0 on success, negative MMLA_E* on failure; mmla_last_error() returns a
 *     thread-local message for the last failing call on the calling thread.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     MMLA_ECUDA.
 */
#ifndef MMLA_B200_H
#define MMLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMLA_OK        0
#define MMLA_EINVAL   -1   /* bad argument                         */
#define MMLA_ECUDA    -2   /* CUDA runtime error / no device       */
#define MMLA_ENOMEM   -3   /* allocation failed                    */
#define MMLA_EUNSUP   -4   /* parameter combination not supported  */

#define MMLA_WINDOW_RECT    0   /* python_speech_features default (what the reference uses) */
#define MMLA_WINDOW_HANN    1
#define MMLA_WINDOW_HAMMING 2

const char* mmla_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int mmla_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
int64_t mmla_launch_count(void);
/* CRC-32C (Castagnoli) of a HOST buffer, as stored in TF tensor-bundle entries. Host only. */
uint32_t mmla_crc32c_host(const void* data_host, size_t n);

/* Launch trace (measurement aid for bench.py's per-kernel roofline): between mmla_trace_begin(stream) and
 * mmla_trace_end every kernel this library launches on `stream` is followed by a CUDA event.  mmla_trace_end
 * waits for them and returns the number of records written: names_host receives the kernel names, one per
 * line ('\n'-separated, NUL-terminated), ms_host[i] the device time between the previous event and record i
 * (kernels of one stream run in order, so that is kernel i's duration plus any launch gap).  Negative = error. */
int mmla_trace_begin(void* stream);
int mmla_trace_end(char* names_host, int64_t names_bytes, float* ms_host, int32_t max_records);

/* Diagnostics of the tensor-core MFCC kernel (both DEVICE pointers, NULL disables each):
 *   dev_buffer  receives the raw stage-1 / stage-2 accumulators, per 64-frame tile 2 x 128 x 256 float32;
 *   dev_stamps  receives clock64 stamps of CTA 0's warp roles, 64 tiles x 32 int64 slots. */
void mmla_debug_mfcc_tc_dump(float* dev_buffer, long long* dev_stamps);
/* Diagnostics of the fused ResNet-stage kernel: dev_stamps (DEVICE pointer, 3 x 64 int64, NULL = off) receives a
 * clock64 timeline of the warp roles of CTA `cta` of each of the three stage launches (row = stage). */
void mmla_debug_resstage_stamps(long long* dev_stamps, int32_t cta);
/* Same for the fused BiLSTM kernel: dev_stamps (DEVICE pointer, T x 16 int64, NULL = off) receives, per time step,
 * the clock64 timeline of CTA `cta` of the forward direction. */
void mmla_debug_lstm_stamps(long long* dev_stamps, int32_t cta);
/* Same for conv_slab_kernel: dev_stamps (DEVICE pointer, 64 x 16 int64, NULL = off) receives, for each of the next 64
 * launches (row = launch ordinal), the clock64 timeline of a mid-image CTA of image `image`:
 * 0 start, 1 set-up done, 2 slab written, 3 slab barrier, 4 first weight chunk landed, 5 MMAs issued, 6 accumulators
 * complete, 7 epilogue done (warp 2), 8 all warps done. */
void mmla_debug_conv_slab_stamps(long long* dev_stamps, int32_t image);
/* One stride-1 'same' Conv2D layer of the classifiers, for layer-level parity tests of the convolution kernels:
 *   y[B,H,W,N] = conv(act(x * pre_scale + pre_shift)) + bias (+ res), x [B,H,W,Cin] float32 NHWC (DEVICE),
 *   w_host [kh*kw*Cin][N] (HOST, Keras HWIO flattened), bias / pre_scale / pre_shift / res / y DEVICE pointers
 *   (pre_scale = NULL: no BN / activation prologue; res = NULL: no residual), pre_act 0 none / 1 ReLU / 2 ELU.
 *   kernel: 0 = conv_igemm_kernel (fp32 CUDA cores), 1 = conv_tc_kernel (tcgen05, im2col gather),
 *           2 = conv_slab_kernel (tcgen05, tap-shifted slab; MMLA_EUNSUP if the layer is not eligible).
 * Synchronous on `stream`. */
int mmla_debug_conv2d(const float* x, const float* w_host, const float* bias, const float* pre_scale, const float* pre_shift,
                      int32_t pre_act, const float* res, float* y, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t N,
                      int32_t kh, int32_t kw, int32_t kernel, void* stream);
/* One res_block conv pair of the overlap classifier (overlap_detector_temp.py:253-280) through resblock2d_fused_kernel:
 *   y[B,H,W,C] = Conv2D(C,(4,1),'same')(ELU(BN2(Conv2D(C,3,'same')(ELU(BN1(x)))))) (+ res), x [B,H,W,Cin] float32 NHWC (DEVICE),
 *   w1_host [9*Cin][C], w2_host [4*C][C] (HOST, Keras HWIO flattened), b1 / b2 / folded BN scale+shift / res / y DEVICE pointers
 *   (res = NULL: no residual; rows of C floats otherwise).  hpool = 1 (pooled blocks: even H, no residual): y is [B,H/2,W,C],
 *   the maximum over the row pairs (2i, 2i+1) of the block output, taken in the kernel's epilogue.
 *   MMLA_EUNSUP if the block is not eligible.  Synchronous on `stream`. */
int mmla_debug_resblock2d(const float* x, const float* w1_host, const float* b1, const float* bn1_scale, const float* bn1_shift,
                          const float* w2_host, const float* b2, const float* bn2_scale, const float* bn2_shift, const float* res,
                          float* y, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t C, int32_t hpool, void* stream);
/* clock64 timeline of resblock2d_fused_kernel: dev_stamps (DEVICE pointer, 16 x 16 int64, NULL = off) receives, for each of the
 * next 16 launches, the stamps of a mid-image CTA of image `image`: 0 start, 1 set-up done, 2 x slab written, 3 slab barrier,
 * 4 conv1 issued, 5 conv1 accumulators complete, 6 u slab written, 7 conv2 issued, 8 conv2 accumulators complete,
 * 9 epilogue done (warp 2), 10 all warps done. */
void mmla_debug_resblock2d_stamps(long long* dev_stamps, int32_t image);
/* Same for resblock2d_persist_kernel: dev_stamps (DEVICE pointer, 4 x 4 x 16 int64, NULL = off) receives, for each of the next
 * four launches, the stamps of work items 4..7 of a mid-grid CTA (16 slots per item): 0/1 fill start / end, 2/3 conv1 issue,
 * 4/5 conv2 issue, 6/7/8 epilogue 1 (accumulator ready, slab free, done), 9/10/11 epilogue 2 (accumulator ready, drained, stored). */
void mmla_debug_resblock2d_persist_stamps(long long* dev_stamps);

/* ------------------------------------------------------------------------------------------
 * Speaker-ID features.
 * Replaces python_speech_features.mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512)
 *   SpeakerIdentification/scripts/speaker_identification.py:89,285,341,386
 *   SpeakerIdentification/scripts/speaker_identification_post_processing.py:256
 * and, when with_deltas=1, the reference's delta(feat,2) twice + concatenate
 *   speaker_identification.py:141-151,387-389
 * and, when pad_frames>0, the zero-pad / truncate to 256 rows of input_feature_gen
 *   speaker_identification.py:391-395  (or to a multiple of 256 rows for whole files, :347-349)
 * ---------------------------------------------------------------------------------------- */
typedef struct MmlaMfccParams {
    int32_t samplerate;     /* 16000 */
    int32_t frame_len;      /* 400  = round_half_up(winlen*samplerate); must be <= nfft */
    int32_t frame_step;     /* 160  = round_half_up(winstep*samplerate) */
    int32_t nfft;           /* 512 (both DFT kernels are sized for exactly 512) */
    int32_t nfilt;          /* 26 = reference; 40 = BASELINE config 3; <= 64 */
    int32_t numcep;         /* 13; <= 14 and <= nfilt */
    int32_t ceplifter;      /* 22 */
    int32_t append_energy;  /* 1: c0 := ln(frame energy) */
    int32_t window;         /* MMLA_WINDOW_* */
    int32_t with_deltas;    /* 0: rows are [numcep]; 1: rows are [numcep | delta | delta-delta] */
    int32_t pad_frames;     /* 0: write T rows per clip; >0: write exactly pad_frames rows
                               (first min(T,pad_frames) real, remainder zero) */
    float   preemph;        /* 0.97 */
    float   lowfreq;        /* 0 */
    float   highfreq;       /* samplerate/2 */
} MmlaMfccParams;

/* Number of frames python_speech_features produces for a clip of n samples. */
int32_t mmla_psf_num_frames(int64_t n_samples, const MmlaMfccParams* p);

/*
 * pcm            int16 mono samples at int16 scale (as scipy.io.wavfile.read returns them),
 *                16-byte aligned, readable up to the next 16-byte boundary past the last sample.
 * clip_off_host  [n_clips] start sample of each clip in `pcm`, or NULL for uniform clips at
 *                clip i -> i*clip_stride.
 * clip_len_host  [n_clips] samples per clip, or NULL for uniform length `clip_len`.
 * out            float32 [n_clips][rows][dim] with clip i at out + i*out_clip_stride floats;
 *                dim = numcep*(with_deltas?3:1); rows = pad_frames>0 ? pad_frames : T(clip).
 */
int mmla_psf_mfcc(const int16_t* pcm, int64_t pcm_total_samples,
                  const int64_t* clip_off_host, const int32_t* clip_len_host,
                  int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                  const MmlaMfccParams* p,
                  float* out, int64_t out_clip_stride, void* stream);

/* Same as mmla_psf_mfcc with rows `out_row_stride` floats apart (>= dim; 0 = dense).  Columns [dim, out_row_stride)
 * of every row are written as zeros: out_row_stride = 40 with with_deltas = 1 is the channel-padded [256,40] layout the
 * speaker classifier's tensor-core stem reads directly (MMLA_INPUT_F32_PAD40), saving a pad pass. */
int mmla_psf_mfcc_rows(const int16_t* pcm, int64_t pcm_total_samples,
                       const int64_t* clip_off_host, const int32_t* clip_len_host,
                       int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                       const MmlaMfccParams* p,
                       float* out, int64_t out_clip_stride, int32_t out_row_stride, void* stream);

/* Replaces delta(feat, N): SpeakerIdentification/scripts/speaker_identification.py:141-151.
 * feat/out float32 [n_frames][dim]; out[t] = sum_{k=-N..N} k*feat[clamp(t+k)] / (2*sum k^2). */
int mmla_delta(const float* feat, int64_t n_frames, int32_t dim, int32_t N, float* out, void* stream);

/* Cepstral mean (variance = 0) or mean-and-variance (variance = 1) normalisation over the frames of each clip, in place:
 * column c of clip i becomes (x - mean_t x) [/ std_t x] over rows [0, n_rows(i)) (population std; a constant column is
 * only centred).  BASELINE.json north_star (3) lists CMVN; the reference applies none (its features are the raw
 * MFCC + delta + delta-delta of speaker_identification.py:386-389), so this is an option, off by default.
 * feat float32, clip i at feat + i*clip_stride floats, rows row_stride floats apart, columns [0, dim) are normalised
 * (dim <= 64); n_rows_dev: DEVICE int32 [n_clips] real rows per clip, or NULL for the uniform n_rows. */
int mmla_cmvn(float* feat, int64_t n_clips, int64_t clip_stride, int32_t row_stride, int32_t dim,
              int32_t n_rows, const int32_t* n_rows_dev, int32_t variance, void* stream);

/* ------------------------------------------------------------------------------------------
 * Overlap-detection features.
 * Replaces OverlapFeaturesGenerator.generate_mels / generate_zcr / generate_zcr_image +
 * plt.imsave(origin='lower') + tf.image.decode_png(.,3)
 *   OverlapDetection/scripts/overlap_features_generator.py:65-151
 *   OverlapDetection/scripts/record_on_pc.py:139,156-158
 * Every clip is zero-padded / truncated to hop*150 = 24000 samples (…generator.py:73-80).
 * Any of the four outputs may be NULL.
 *   s_db      float32 [n_clips][n_mels][151]   power_to_db(ref=max, top_db=80)
 *   s_db_norm float32 [n_clips][n_mels][151]   min-max normalised
 *   zcr       float32 [n_clips][151]
 *   image     uint8   [n_clips][n_mels][151][3] row 0 = highest mel band, trunc(v*255)
 * ---------------------------------------------------------------------------------------- */
int mmla_overlap_features(const int16_t* pcm, int64_t pcm_total_samples,
                          const int64_t* clip_off_host, const int32_t* clip_len_host,
                          int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                          int32_t n_mels,
                          float* s_db, float* s_db_norm, float* zcr, uint8_t* image,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Classifier forward pass.
 * Replaces tf.keras.models.load_model(dir) / model.predict(x)
 *   OverlapDetection/scripts/record_on_pc.py:87-88,159
 *   SpeakerIdentification/scripts/record_on_pc.py:76-77,136
 * Weights are passed as one float32 HOST blob laid out by mmla_audio_b200/models.py
 * (TF layouts kept: HWIO / WIO kernels, LSTM [in,4u] i,f,c,o, Dense [in,out]).
 * ---------------------------------------------------------------------------------------- */
#define MMLA_NET_OVERLAP 0   /* x: uint8 [B,128,151,3] (or float32 0..255)  -> prob [B,2]   */
#define MMLA_NET_SPEAKER 1   /* x: float32 [B,256,39]                       -> prob [B,n]   */
#define MMLA_HEAD_SOFTMAX 0
#define MMLA_HEAD_SIGMOID 1
/* Arithmetic of the conv / LSTM matrix products (features are always fp32):
 *   FP32: CUDA-core implicit GEMM, fp32 operands and accumulation (bit-faithful layer semantics);
 *   TF32: tcgen05 tensor cores, TF32 operands (10-bit mantissa), fp32 accumulation in TMEM. */
#define MMLA_PRECISION_FP32 0
#define MMLA_PRECISION_TF32 1
/*   F16: as TF32, except that the overlap net's res_block conv pairs (99 % of its FLOP) and both nets' LSTM recurrence take
 *        fp16 operands (`kind::f16`: the same 11 significant bits as TF32, half the bytes; conversions saturate at 65504;
 *        the LSTM's h lies in (-1, 1)), fp32 accumulation.  Every tensor in HBM stays fp32. */
#define MMLA_PRECISION_F16 2

typedef struct MmlaNet MmlaNet;   /* opaque */

int mmla_net_create(int32_t kind, int32_t n_classes, int32_t head_activation,
                    const float* weights_host, int64_t n_weights, MmlaNet** out_net);
void mmla_net_destroy(MmlaNet* net);
int mmla_net_set_precision(MmlaNet* net, int32_t mode);
/* Bytes of device workspace needed for a forward pass over `batch` clips. */
int64_t mmla_net_workspace_bytes(const MmlaNet* net, int64_t batch);
#define MMLA_INPUT_F32       0   /* float32 in the net's own input shape                                   */
#define MMLA_INPUT_U8        1   /* uint8 image tensor (overlap net only)                                  */
#define MMLA_INPUT_F32_PAD40 2   /* speaker net, TF32 mode: float32 [B,256,40], channel 39 = 0             */
/*
 * x_is_u8: one of MMLA_INPUT_* (historical name: 1 when x is the uint8 image tensor, 0 for float32 input).
 * prob    float32 [batch][n_classes]; labels int32 [batch] = argmax (first max wins), may be NULL.
 */
int mmla_net_forward(MmlaNet* net, const void* x, int32_t x_is_u8, int64_t batch,
                     void* workspace, int64_t workspace_bytes,
                     float* prob, int32_t* labels, void* stream);

/*
 * Speaker net, TF32 mode, straight from the MFCC-13 rows (the label pipeline of
 * SpeakerIdentification/scripts/record_on_pc.py:120-137 when the [256,39] feature tensor itself is not wanted):
 *   cepstra   float32, clip i at cepstra + i*cep_clip_stride floats, rows of 16 floats (13 cepstra + 3 ignored),
 *             n_frames rows per clip (the psf frame count, 1..256) — what mmla_psf_mfcc_rows(with_deltas=0,
 *             pad_frames=0, out_row_stride=16) writes.
 * The stem kernel builds delta / delta-delta (speaker_identification.py:141-151,387-389) and the zero rows up to 256
 * (:391-395) on the fly, so the feature tensor never exists in HBM.  Results equal mmla_net_forward on the features.
 */
int mmla_net_forward_cepstra(MmlaNet* net, const float* cepstra, int64_t cep_clip_stride, int32_t n_frames, int64_t batch,
                             void* workspace, int64_t workspace_bytes, float* prob, int32_t* labels, void* stream);

/* ------------------------------------------------------------------------------------------
 * Enrollment: the frozen trunk's 512-d output and the fit of the transfer head.
 * Replaces the first phase of transfer_learning, SpeakerIdentification/scripts/speaker_identification.py:401-432:
 *   sliced_base_model = Model(base_model.input, base_model.layers[-2].output)          -> mmla_net_embed
 *   Dense(dim, activation='sigmoid', name='customized_dense'); compile(loss="categorical_crossentropy",
 *   optimizer=RMSprop(lr=0.0001)); fit(batch_size=16, epochs=500)                        -> mmla_head_fit
 * mmla_net_embed: like mmla_net_forward (same x / x_is_u8 / workspace), writes embed float32 [batch][512] =
 *   [forward h | backward h] of the Bidirectional LSTM (the input of the Dense head).
 * mmla_head_fit: embed [n_samples][512], y_onehot float32 [n_samples][n_classes] (n_classes <= 64), order int32
 *   [epochs][n_samples] = the sample visiting order of every epoch (Keras reshuffles per epoch; supplied by the caller so
 *   a fit is reproducible), batch_size <= 32; kernel [512][n_classes] and bias [n_classes] (DEVICE) hold the initial
 *   weights on entry and the fitted ones on return; loss_out [epochs] (DEVICE, may be NULL) the mean training loss per
 *   epoch.  Loss and optimiser as Keras defines them (see csrc/head_fit.cu).  All pointers DEVICE.
 * ---------------------------------------------------------------------------------------- */
int mmla_net_embed(MmlaNet* net, const void* x, int32_t x_is_u8, int64_t batch, void* workspace, int64_t workspace_bytes,
                   float* embed, void* stream);
int mmla_head_fit(const float* embed, const float* y_onehot, int64_t n_samples, int32_t n_classes, const int32_t* order,
                  int32_t epochs, int32_t batch_size, float lr, float rho, float eps, float* kernel, float* bias,
                  float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stationary spectral-gating noise reduction.
 * Replaces  nr.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True) + sf.write(filepath, ., 16000)
 *   OverlapDetection/scripts/record_on_pc.py:208-212; overlap_detection_post_processing.py:128-133 (and the
 *   SpeakerIdentification copies) with noisereduce's defaults: n_fft 1024, hop 256, n_std_thresh_stationary 1.5,
 *   prop_decrease 1, 500 Hz / 50 ms mask smoothing, 30000-sample chunk padding; output = PCM_16 as libsndfile writes it.
 * mmla_noise_profile: noise int16 [n_samples] (the ambient-noise recording, first 600000 samples used) ->
 *   thresh_out float32 [513] (DEVICE) = per-bin mean + n_std_thresh * std of the noise spectrogram in dB.
 * mmla_noise_gate: pcm int16 [n_clips][clip_stride], clip_len_dev DEVICE int32 [n_clips] or NULL (uniform clip_len,
 *   <= 600000) -> out int16 [n_clips][out_stride], samples [0, len) of every clip (the rest untouched).
 * ---------------------------------------------------------------------------------------- */
int mmla_noise_profile(const int16_t* noise, int64_t n_samples, float n_std_thresh, float* thresh_out, void* stream);
int mmla_noise_gate(const int16_t* pcm, int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                    const int32_t* clip_len_dev, const float* thresh, int16_t* out, int64_t out_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * Silence removal: WebRTC VAD, aggressiveness 3, 16 kHz, 30 ms frames + vad_collector.
 * Replaces  vad = webrtcvad.Vad(3); vad.is_speech(frame.bytes, sample_rate)
 *           frame_generator(30, audio, sample_rate); vad_collector(sample_rate, 30, 300, vad, frames)
 *           and the rewrite of the clip from the yielded segments in save_wave_file(..., silence_remove=True)
 *   OverlapDetection/scripts/record_on_pc.py:33,214-226,229-295 (same code in SpeakerIdentification/scripts/
 *   record_on_pc.py:207-273 and both *_post_processing.py files), whose consequence `len(sig) < 4000 => 'silent'`
 *   (record_on_pc.py:142; speaker_identification.py:375) is applied per clip by the pipelines.
 * pcm            int16 [n_clips][clip_stride] (DEVICE), clips starting on 8-byte boundaries.
 * clip_len_dev   DEVICE int32 [n_clips] samples per clip, or NULL for the uniform clip_len.
 * clips_per_stream  the detector adapts from frame to frame and, in the reference (one module-global Vad object), from
 *                clip to clip: every run of this many consecutive clips is one sequential stream starting from a fresh
 *                detector.  1 = independent clips (batch mode, one thread per clip); n_clips = one session in order.
 * Outputs (DEVICE; each may be NULL, pcm_out needs keep):
 *   speech       uint8 [n_clips][max_frames]  is_speech of frame f (0 past the clip's frame count)
 *   keep         uint8 [n_clips][max_frames]  1 where vad_collector yields the frame
 *   voiced_len   int32 [n_clips]              480 * kept frames = the rewritten clip's length in samples
 *   pcm_out      int16 [n_clips][out_stride]  kept frames, concatenated in order (samples past voiced_len untouched)
 * mmla_vad_num_frames: frames frame_generator yields for n samples (`while offset + n < len(audio)`, byte offsets).
 * ---------------------------------------------------------------------------------------- */
int32_t mmla_vad_num_frames(int32_t n_samples);
int mmla_vad_trim(const int16_t* pcm, int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                  const int32_t* clip_len_dev, int64_t clips_per_stream,
                  uint8_t* speech, uint8_t* keep, int32_t max_frames, int32_t* voiced_len,
                  int16_t* pcm_out, int64_t out_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * Label tallies.  Replaces the counting loops of
 *   OverlapDetection/scripts/overlap_degree_distribution.py:49-61
 *   SpeakerIdentification/scripts/speaker_time_distribution.py:52-80
 * labels int32 [n]; entries outside [0,n_classes) (e.g. the 'silent' sentinel -1) are counted
 * in counts[n_classes].  counts int64 [n_classes+1] is ADDED to (zero it first).
 * ---------------------------------------------------------------------------------------- */
int mmla_tally(const int32_t* labels, int64_t n, int32_t n_classes, int64_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Synthetic PCM (stands in for PyAudio capture, OverlapDetection/scripts/record_on_pc.py:115-124).
 * Integer-only generator: clip i depends only on (seed, first_clip+i); bit-identical to
 * oracle/synth.py.  sine_table is int16[1024] on the device.
 * ---------------------------------------------------------------------------------------- */
int mmla_synth_pcm(int16_t* pcm, int64_t first_clip, int64_t n_clips, int32_t clip_len,
                   int64_t clip_stride, uint32_t seed, const int16_t* sine_table, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMLA_B200_H */
