// Voice-activity trimming on the device: WebRTC's fixed-point VAD (aggressiveness 3, 16 kHz, 30 ms frames) + the
// reference's `vad_collector` hysteresis + the rewrite of the clip from the collected frames.
//
// Replaces, for batches of clips resident in HBM:
//   vad = webrtcvad.Vad(3); vad.is_speech(frame.bytes, sample_rate)    OverlapDetection/scripts/record_on_pc.py:33,254
//   frame_generator(30, audio, sample_rate)                            record_on_pc.py:229-244
//   vad_collector(sample_rate, 30, 300, vad, frames)                   record_on_pc.py:247-295
//   save_wave_file(..., silence_remove=True): clip := b''.join(segments)   record_on_pc.py:214-226
//   (same code in SpeakerIdentification/scripts/record_on_pc.py:207-273 and both *_post_processing.py files)
// whose consequence — `len(sig) < 4000` => 'silent' (record_on_pc.py:142, speaker_identification.py:375) — decides the
// third label of both tallies.
//
// The detector is a chain of integer IIR filters (two all-pass branches per band split, a six-band tree) feeding a
// two-Gaussian noise / speech model per band that adapts on every frame; every state carries from frame to frame, and in
// the reference — one module-global Vad object — from clip to clip.  The recurrences round in 16 / 32-bit fixed point, so
// they cannot be re-associated into a scan: a STREAM of clips is inherently sequential.  Parallelism is across streams:
// one thread owns one stream (= `clips_per_stream` consecutive clips; 1 = every clip starts from a fresh detector, the
// batch mode; n_clips = one whole session in the reference's order).  All arithmetic is integer; results are bit-exact
// against the CPU restatement (tests/test_vad_gpu.py).  The working set per thread (two 120- and two 60-sample band
// buffers, 6 x 16 minimum trackers) lives in local memory, which the hardware interleaves by lane, i.e. every access of a
// warp in lockstep is one coalesced L1 line.
#include "common.cuh"

namespace {

constexpr int kBands = 6;
constexpr int kFrame = 480;                 // 30 ms at 16 kHz
constexpr int kRing = 10;                   // vad_collector: padding_duration_ms / frame_duration_ms = 300 / 30

// start values of the band models, Q7 (noise / speech; index = band + 6 * gaussian)
__constant__ int16_t cNoiseW[12] = {34, 62, 72, 66, 53, 25, 94, 66, 56, 62, 75, 103};
__constant__ int16_t cSpeechW[12] = {48, 82, 45, 87, 50, 47, 80, 46, 83, 41, 78, 81};
__constant__ int16_t cNoiseMean0[12] = {6738, 4892, 7065, 6715, 6771, 3369, 7646, 3863, 7820, 7266, 5020, 4362};
__constant__ int16_t cSpeechMean0[12] = {8306, 10085, 10078, 11823, 11843, 6309, 9473, 9571, 10879, 7581, 8180, 7483};
__constant__ int16_t cNoiseStd0[12] = {378, 1064, 493, 582, 688, 593, 474, 697, 475, 688, 421, 455};
__constant__ int16_t cSpeechStd0[12] = {555, 505, 567, 524, 585, 1231, 509, 828, 492, 1540, 1079, 850};
__constant__ int16_t cMinDiff[kBands] = {544, 544, 576, 576, 576, 576};
__constant__ int16_t cMaxSpeech[kBands] = {11392, 11392, 11520, 11520, 11520, 11520};
__constant__ int16_t cMaxNoise[kBands] = {9216, 9088, 8960, 8832, 8704, 8576};
__constant__ int16_t cBandOffset[kBands] = {368, 368, 272, 176, 176, 176};

// mode 3, 30 ms frames
constexpr int kHang1 = 2, kHang2 = 3, kLocalThr = 94, kGlobalThr = 1100;

struct Detector {
    int32_t ds[2];                 // 16 -> 8 kHz all-pass memories
    int16_t up[5], lo[5];          // band-split all-pass memories
    int16_t hp[4];                 // 80 Hz high-pass memory
    int16_t nmean[12], smean[12], nstd[12], sstd[12];
    int16_t low[kBands * 16], age[kBands * 16];
    int16_t median[kBands];
    int32_t frames;
    int16_t hang, run;
};

__device__ __forceinline__ int16_t s16(int v) { return static_cast<int16_t>(v); }
__device__ __forceinline__ int norm_w32(int32_t a) {
    if (a == 0) return 0;
    if (a < 0) a = ~a;
    return __clz(a) - 1;
}
__device__ __forceinline__ int32_t div_w32_w16(int32_t num, int16_t den) { return den != 0 ? num / den : 0x7FFFFFFF; }

__device__ void detector_reset(Detector& d) {
    d.ds[0] = d.ds[1] = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) d.up[i] = d.lo[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) d.hp[i] = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        d.nmean[i] = cNoiseMean0[i];
        d.smean[i] = cSpeechMean0[i];
        d.nstd[i] = cNoiseStd0[i];
        d.sstd[i] = cSpeechStd0[i];
    }
    for (int i = 0; i < kBands * 16; ++i) {
        d.low[i] = 10000;
        d.age[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < kBands; ++i) d.median[i] = 1600;
    d.frames = 0;
    d.hang = d.run = 0;
}

// One band split: two first-order all-pass branches over the even / odd samples, then sum and difference.
// Both branches advance in the same loop (two independent dependency chains).
__device__ __forceinline__ void band_split(const int16_t* in, int half, int16_t& up_state, int16_t& lo_state, int16_t* hp,
                                           int16_t* lp) {
    int32_t su = static_cast<int32_t>(up_state) << 16, sl = static_cast<int32_t>(lo_state) << 16;
    for (int i = 0; i < half; ++i) {
        const int a = in[2 * i], b = in[2 * i + 1];
        const int16_t yu = s16((su + 20972 * a) >> 16);
        su = static_cast<int32_t>(static_cast<uint32_t>((a << 14) - 20972 * yu) << 1);
        const int16_t yl = s16((sl + 5571 * b) >> 16);
        sl = static_cast<int32_t>(static_cast<uint32_t>((b << 14) - 5571 * yl) << 1);
        hp[i] = s16(yu - yl);
        lp[i] = s16(yl + yu);
    }
    up_state = s16(su >> 16);
    lo_state = s16(sl >> 16);
}

// 10 log10(energy) of a band in Q4 (+ the band's offset); also feeds the frame's coarse power indicator.
__device__ __forceinline__ int16_t band_log_energy(const int16_t* v, int n, int16_t offset, int16_t& total) {
    int16_t smax = -1;
    for (int i = 0; i < n; ++i) {
        const int16_t a = s16(v[i] > 0 ? v[i] : -v[i]);                  // -(-32768) wraps, as upstream
        smax = a > smax ? a : smax;
    }
    int scaling = 0;
    if (smax != 0) {
        const int nbits = 32 - __clz(n), t = norm_w32(static_cast<int32_t>(smax) * smax);
        scaling = t > nbits ? 0 : nbits - t;
    }
    int32_t en = 0;
    for (int i = 0; i < n; ++i) en += (static_cast<int32_t>(v[i]) * v[i]) >> scaling;
    uint32_t energy = static_cast<uint32_t>(en);
    if (energy == 0) return offset;
    int rshifts = scaling;
    const int nrm = 17 - __clz(energy);
    rshifts += nrm;
    energy = nrm < 0 ? energy << -nrm : energy >> nrm;
    const int16_t log2e = s16(14336 + s16((energy & 0x3FFFu) >> 4));
    int16_t le = s16(((24660 * log2e) >> 19) + ((rshifts * 24660) >> 9));
    if (le < 0) le = 0;
    le = s16(le + offset);
    if (total <= 10) {
        if (rshifts >= 0) total = s16(total + 11);
        else total = s16(total + s16(energy >> -rshifts));
    }
    return le;
}

// (1/s) exp(-(x-m)^2 / 2 s^2) in Q20 and delta = (x - m) / s^2 in Q11
__device__ __forceinline__ int32_t gauss(int16_t x, int16_t mean, int16_t sd, int16_t& delta) {
    const int16_t inv = s16(div_w32_w16(131072 + (sd >> 1), sd));
    const int16_t q = s16(inv >> 2);
    const int16_t inv2 = s16((q * q) >> 2);
    const int16_t dx = s16(s16(x << 3) - mean);
    delta = s16((inv2 * dx) >> 10);
    const int32_t e = (delta * dx) >> 9;
    int16_t ev = 0;
    if (e < 22005) {
        int16_t t = s16(-s16((5909 * e) >> 12));
        ev = s16(0x0400 | (t & 0x03FF));
        t = s16(t ^ 0xFFFF);
        t = s16(t >> 10);
        t = s16(t + 1);
        ev = s16(ev >> t);
    }
    return inv * ev;
}

// smoothed minimum of a band's feature over the last 100 frames (16 smallest values with their ages)
__device__ int16_t track_minimum(Detector& d, int16_t x, int band) {
    int16_t* low = &d.low[band * 16];
    int16_t* age = &d.age[band * 16];
    for (int i = 0; i < 16; ++i) {
        if (age[i] != 100) {
            age[i]++;
        } else {
            for (int j = i; j < 15; ++j) {
                low[j] = low[j + 1];
                age[j] = age[j + 1];
            }
            age[15] = 101;
            low[15] = 10000;
        }
    }
    // the 16 values stay sorted ascending, so the insertion slot is the number of entries <= x
    int pos = 0;
    for (int i = 0; i < 16; ++i) pos += (low[i] <= x) ? 1 : 0;
    if (pos < 16) {
        for (int i = 15; i > pos; --i) {
            low[i] = low[i - 1];
            age[i] = age[i - 1];
        }
        low[pos] = x;
        age[pos] = 1;
    }
    int16_t cur = 1600, alpha = 0;
    if (d.frames > 2) cur = low[2];
    else if (d.frames > 0) cur = low[0];
    if (d.frames > 0) alpha = cur < d.median[band] ? 6553 : 32439;
    int32_t t = (alpha + 1) * d.median[band];
    t += (32767 - alpha) * cur;
    t += 16384;
    d.median[band] = s16(t >> 15);
    return d.median[band];
}

__device__ __forceinline__ int32_t pair_average(int16_t* m, int band, int16_t shift, const int16_t* w) {
    m[band] = s16(m[band] + shift);
    m[band + 6] = s16(m[band + 6] + shift);
    return m[band] * w[band] + m[band + 6] * w[band + 6];
}

// likelihood-ratio decision + model adaptation for one frame; returns the raw flag (0, 1, or 2 + hang-over)
__device__ int decide_and_adapt(Detector& d, const int16_t (&feat)[kBands], int16_t power) {
    int flag = 0;
    if (power > 10) {
        int16_t dN[12], dS[12], gN[12], gS[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) gN[i] = gS[i] = 0;
        int32_t llr_sum = 0;
#pragma unroll
        for (int c = 0; c < kBands; ++c) {
            int32_t pn[2], ps[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int g = c + 6 * k;
                pn[k] = cNoiseW[g] * gauss(feat[c], d.nmean[g], d.nstd[g], dN[g]);
                ps[k] = cSpeechW[g] * gauss(feat[c], d.smean[g], d.sstd[g], dS[g]);
            }
            const int32_t h0 = pn[0] + pn[1], h1 = ps[0] + ps[1];
            const int sh0 = h0 == 0 ? 31 : norm_w32(h0), sh1 = h1 == 0 ? 31 : norm_w32(h1);
            const int16_t llr = s16(sh0 - sh1);
            llr_sum += llr * (6 + 2 * c);                                  // spectrum weights 6, 8, .., 16
            if (llr * 4 > kLocalThr) flag = 1;
            const int16_t q0 = s16(h0 >> 12);
            if (q0 > 0) {
                gN[c] = s16(div_w32_w16(static_cast<int32_t>((static_cast<uint32_t>(pn[0]) & 0xFFFFF000u) << 2), q0));
                gN[c + 6] = s16(16384 - gN[c]);
            } else {
                gN[c] = 16384;
            }
            const int16_t q1 = s16(h1 >> 12);
            if (q1 > 0) {
                gS[c] = s16(div_w32_w16(static_cast<int32_t>((static_cast<uint32_t>(ps[0]) & 0xFFFFF000u) << 2), q1));
                gS[c + 6] = s16(16384 - gS[c]);
            }
        }
        if (llr_sum >= kGlobalThr) flag |= 1;

        int16_t maxspe = 12800;
#pragma unroll
        for (int c = 0; c < kBands; ++c) {
            const int16_t fmin = track_minimum(d, feat[c], c);
            const int16_t nglob8 = s16(pair_average(d.nmean, c, 0, cNoiseW) >> 6);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int g = c + 6 * k;
                const int16_t nmk = d.nmean[g], smk = d.smean[g];
                int16_t nsk = d.nstd[g], ssk = d.sstd[g];
                int16_t nmk2 = nmk;
                if (!flag) {
                    const int16_t dl = s16((gN[g] * dN[g]) >> 11);
                    nmk2 = s16(nmk + s16((dl * 655) >> 22));
                }
                const int16_t nd = s16((fmin << 4) - nglob8);
                int16_t nmk3 = s16(nmk2 + s16((nd * 154) >> 9));
                const int16_t floor_q7 = s16((k + 5) << 7), ceil_q7 = s16((72 + k - c) << 7);
                if (nmk3 < floor_q7) nmk3 = floor_q7;
                if (nmk3 > ceil_q7) nmk3 = ceil_q7;
                d.nmean[g] = nmk3;
                if (flag) {
                    const int16_t dl = s16((gS[g] * dS[g]) >> 11);
                    int16_t t = s16((dl * 6554) >> 21);
                    int16_t smk2 = s16(smk + ((t + 1) >> 1));
                    const int16_t lo_lim = k == 0 ? 640 : 768, hi_lim = s16(maxspe + 640);
                    if (smk2 < lo_lim) smk2 = lo_lim;
                    if (smk2 > hi_lim) smk2 = hi_lim;
                    d.smean[g] = smk2;
                    t = s16(feat[c] - s16((smk + 4) >> 3));
                    int32_t a = ((dS[g] * t) >> 3) - 4096;
                    t = s16(gS[g] >> 2);
                    a = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(t)) * static_cast<uint32_t>(a)) >> 4;
                    const int16_t den = s16(ssk * 10);
                    t = a > 0 ? s16(div_w32_w16(a, den)) : s16(-s16(div_w32_w16(-a, den)));
                    t = s16(t + 128);
                    ssk = s16(ssk + (t >> 8));
                    if (ssk < 384) ssk = 384;
                    d.sstd[g] = ssk;
                } else {
                    int16_t t = s16(feat[c] - (nmk >> 3));
                    int32_t a = ((dN[g] * t) >> 3) - 4096;
                    t = s16((gN[g] + 2) >> 2);
                    a = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(t)) * static_cast<uint32_t>(a)) >> 14;
                    t = a > 0 ? s16(div_w32_w16(a, nsk)) : s16(-s16(div_w32_w16(-a, nsk)));
                    t = s16(t + 32);
                    nsk = s16(nsk + (t >> 6));
                    if (nsk < 384) nsk = 384;
                    d.nstd[g] = nsk;
                }
            }
            // keep the two models apart and inside their ranges
            int32_t nglob = pair_average(d.nmean, c, 0, cNoiseW);
            int32_t sglob = pair_average(d.smean, c, 0, cSpeechW);
            const int16_t diff = s16(s16(sglob >> 9) - s16(nglob >> 9));
            if (diff < cMinDiff[c]) {
                const int16_t gap = s16(cMinDiff[c] - diff);
                sglob = pair_average(d.smean, c, s16((13 * gap) >> 2), cSpeechW);
                nglob = pair_average(d.nmean, c, s16(-s16((3 * gap) >> 2)), cNoiseW);
            }
            maxspe = cMaxSpeech[c];
            int16_t over = s16(sglob >> 7);
            if (over > maxspe) {
                over = s16(over - maxspe);
                d.smean[c] = s16(d.smean[c] - over);
                d.smean[c + 6] = s16(d.smean[c + 6] - over);
            }
            over = s16(nglob >> 7);
            if (over > cMaxNoise[c]) {
                over = s16(over - cMaxNoise[c]);
                d.nmean[c] = s16(d.nmean[c] - over);
                d.nmean[c + 6] = s16(d.nmean[c + 6] - over);
            }
        }
        d.frames++;
    }
    if (!flag) {
        if (d.hang > 0) {
            flag = 2 + d.hang;
            d.hang--;
        }
        d.run = 0;
    } else {
        d.run++;
        if (d.run > 6) {
            d.run = 6;
            d.hang = kHang2;
        } else {
            d.hang = kHang1;
        }
    }
    return flag;
}

// One 30 ms frame.  The 16 -> 8 kHz decimator (two all-pass branches, one output per input pair) feeds the first band
// split directly: every two decimated samples advance the 0-4 kHz split by one step, so the 240-sample narrow-band frame
// is never stored.
__device__ int frame_is_speech(Detector& d, const int16_t* __restrict__ x) {
    int16_t h120[120], l120[120], h60[60], l60[60];
    {
        int32_t t1 = d.ds[0], t2 = d.ds[1];
        int32_t su = static_cast<int32_t>(d.up[0]) << 16, sl = static_cast<int32_t>(d.lo[0]) << 16;
        const uint2* x4 = reinterpret_cast<const uint2*>(x);                // frames start on 8-byte boundaries (see host)
        for (int i = 0; i < 120; ++i) {
            const uint2 w = __ldg(&x4[i]);
            const int s0 = static_cast<int16_t>(w.x & 0xffffu), s1 = static_cast<int16_t>(w.x >> 16);
            const int s2 = static_cast<int16_t>(w.y & 0xffffu), s3 = static_cast<int16_t>(w.y >> 16);
            int16_t a1 = s16((t1 >> 1) + ((5243 * s0) >> 14));
            t1 = s0 - ((5243 * a1) >> 12);
            int16_t a2 = s16((t2 >> 1) + ((1392 * s1) >> 14));
            t2 = s1 - ((1392 * a2) >> 12);
            const int a = s16(a1 + a2);                                      // narrow-band sample 2 i
            a1 = s16((t1 >> 1) + ((5243 * s2) >> 14));
            t1 = s2 - ((5243 * a1) >> 12);
            a2 = s16((t2 >> 1) + ((1392 * s3) >> 14));
            t2 = s3 - ((1392 * a2) >> 12);
            const int b = s16(a1 + a2);                                      // narrow-band sample 2 i + 1
            const int16_t yu = s16((su + 20972 * a) >> 16);
            su = static_cast<int32_t>(static_cast<uint32_t>((a << 14) - 20972 * yu) << 1);
            const int16_t yl = s16((sl + 5571 * b) >> 16);
            sl = static_cast<int32_t>(static_cast<uint32_t>((b << 14) - 5571 * yl) << 1);
            h120[i] = s16(yu - yl);                                          // 2000 - 4000 Hz
            l120[i] = s16(yl + yu);                                          // 0 - 2000 Hz
        }
        d.ds[0] = t1;
        d.ds[1] = t2;
        d.up[0] = s16(su >> 16);
        d.lo[0] = s16(sl >> 16);
    }
    int16_t feat[kBands], power = 0;
    band_split(h120, 60, d.up[1], d.lo[1], h60, l60);                         // 3000-4000 | 2000-3000
    feat[5] = band_log_energy(h60, 60, cBandOffset[5], power);
    feat[4] = band_log_energy(l60, 60, cBandOffset[4], power);
    band_split(l120, 60, d.up[2], d.lo[2], h60, l60);                         // 1000-2000 | 0-1000
    feat[3] = band_log_energy(h60, 60, cBandOffset[3], power);
    band_split(l60, 30, d.up[3], d.lo[3], h120, l120);                        // 500-1000 | 0-500
    feat[2] = band_log_energy(h120, 30, cBandOffset[2], power);
    band_split(l120, 15, d.up[4], d.lo[4], h60, l60);                         // 250-500 | 0-250
    feat[1] = band_log_energy(h60, 15, cBandOffset[1], power);
    for (int i = 0; i < 15; ++i) {                                            // 80 Hz high-pass of the lowest band
        int32_t t = 6631 * l60[i];
        t += -13262 * d.hp[0];
        t += 6631 * d.hp[1];
        d.hp[1] = d.hp[0];
        d.hp[0] = l60[i];
        t -= -7756 * d.hp[2];
        t -= 5620 * d.hp[3];
        d.hp[3] = d.hp[2];
        d.hp[2] = s16(t >> 14);
        h120[i] = d.hp[2];
    }
    feat[0] = band_log_energy(h120, 15, cBandOffset[0], power);
    return decide_and_adapt(d, feat, power) > 0 ? 1 : 0;
}

// frame_generator's count: `while offset + n < len(audio)` on byte offsets (a clip that is an exact multiple of 480
// samples loses its last frame)
__host__ __device__ __forceinline__ int vad_frames(int n_samples) {
    return n_samples <= kFrame ? 0 : (n_samples - 1) / kFrame;
}

struct VadArgs {
    const int16_t* pcm;
    const int32_t* clip_len_dev;
    uint8_t* speech;
    uint8_t* keep;
    int32_t* voiced_len;
    long long n_clips, clip_stride, clips_per_stream, n_streams;
    int clip_len, max_frames;
};

__global__ void __launch_bounds__(32) vad_decide_kernel(const VadArgs a) {
    const long long stream = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (stream >= a.n_streams) return;
    Detector d;
    detector_reset(d);
    const long long c0 = stream * a.clips_per_stream;
    const long long c1 = min(c0 + a.clips_per_stream, a.n_clips);
    for (long long c = c0; c < c1; ++c) {
        const int len = a.clip_len_dev ? a.clip_len_dev[c] : a.clip_len;
        const int nf = min(vad_frames(len), a.max_frames);
        const int16_t* x = a.pcm + c * a.clip_stride;
        // vad_collector: a 10-frame ring of decisions; enter TRIGGERED when > 90 % of a full ring is voiced (all 10) and
        // collect the ring; leave when > 90 % is unvoiced.  Every frame appended to voiced_frames is eventually yielded.
        uint32_t ring = 0;
        int ring_len = 0, kept = 0;
        bool triggered = false;
        uint8_t* sp = a.speech ? a.speech + c * a.max_frames : nullptr;
        uint8_t* kp = a.keep ? a.keep + c * a.max_frames : nullptr;
        for (int f = 0; f < nf; ++f) {
            const int v = frame_is_speech(d, x + f * kFrame);
            if (sp) sp[f] = static_cast<uint8_t>(v);
            ring = ((ring << 1) | static_cast<uint32_t>(v)) & ((1u << kRing) - 1u);
            ring_len = min(ring_len + 1, kRing);
            const int voiced = __popc(ring & ((1u << ring_len) - 1u));
            if (!triggered) {
                if (kp) kp[f] = 0;
                if (10 * voiced > 9 * kRing) {
                    triggered = true;
                    if (kp)
                        for (int j = f - ring_len + 1; j <= f; ++j) kp[j] = 1;
                    kept += ring_len;
                    ring = 0;
                    ring_len = 0;
                }
            } else {
                if (kp) kp[f] = 1;
                ++kept;
                if (10 * (ring_len - voiced) > 9 * kRing) {
                    triggered = false;
                    ring = 0;
                    ring_len = 0;
                }
            }
        }
        if (kp)
            for (int f = nf; f < a.max_frames; ++f) kp[f] = 0;
        if (sp)
            for (int f = nf; f < a.max_frames; ++f) sp[f] = 0;
        if (a.voiced_len) a.voiced_len[c] = kept * kFrame;
    }
}

// Rewrites each clip as the concatenation of its kept frames (what the reference writes back into the WAV).
// One CTA per clip; the kept frames' output slots come from a prefix sum of the mask.
__global__ void __launch_bounds__(128) vad_compact_kernel(const int16_t* __restrict__ pcm, long long clip_stride,
                                                          const uint8_t* __restrict__ keep, int max_frames,
                                                          int16_t* __restrict__ out, long long out_stride, int vec_ok) {
    __shared__ int slot[1024];
    const long long c = blockIdx.x;
    const uint8_t* kp = keep + c * max_frames;
    if (threadIdx.x == 0) {
        int n = 0;
        for (int f = 0; f < max_frames; ++f) {
            slot[f] = kp[f] ? n : -1;
            n += kp[f] ? 1 : 0;
        }
    }
    __syncthreads();
    const int16_t* src = pcm + c * clip_stride;
    int16_t* dst = out + c * out_stride;
    if (vec_ok) {
        const int per = kFrame / 8;                                       // 60 16-byte words per frame
        for (int i = threadIdx.x; i < max_frames * per; i += blockDim.x) {
            const int f = i / per, w = i - f * per;
            const int s = slot[f];
            if (s >= 0) reinterpret_cast<uint4*>(dst + s * kFrame)[w] = __ldg(reinterpret_cast<const uint4*>(src + f * kFrame) + w);
        }
    } else {
        for (int i = threadIdx.x; i < max_frames * kFrame; i += blockDim.x) {
            const int f = i / kFrame, w = i - f * kFrame;
            const int s = slot[f];
            if (s >= 0) dst[s * kFrame + w] = src[f * kFrame + w];
        }
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int32_t mmla_vad_num_frames(int32_t n_samples) { return vad_frames(n_samples); }

extern "C" __attribute__((visibility("default"))) int mmla_vad_trim(const int16_t* pcm, int64_t n_clips, int32_t clip_len,
                                                                    int64_t clip_stride, const int32_t* clip_len_dev,
                                                                    int64_t clips_per_stream, uint8_t* speech, uint8_t* keep,
                                                                    int32_t max_frames, int32_t* voiced_len, int16_t* pcm_out,
                                                                    int64_t out_stride, void* stream) {
    MMLA_REQUIRE(pcm, MMLA_EINVAL, "vad: null pcm");
    MMLA_REQUIRE(n_clips >= 0 && clip_len >= 0 && clip_stride >= 0 && clips_per_stream >= 1, MMLA_EINVAL, "vad: bad geometry");
    MMLA_REQUIRE(max_frames >= 0 && max_frames <= 1024, MMLA_EINVAL, "vad: max_frames %d outside [0, 1024] (30.7 s per clip)", max_frames);
    MMLA_REQUIRE((clip_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(pcm) & 7) == 0, MMLA_EINVAL,
                 "vad: clips must start on 8-byte boundaries (clip_stride %% 4 == 0)");
    MMLA_REQUIRE(!pcm_out || keep, MMLA_EINVAL, "vad: pcm_out needs the keep mask buffer");
    MMLA_REQUIRE(!pcm_out || out_stride >= static_cast<int64_t>(max_frames) * kFrame || out_stride >= clip_len, MMLA_EINVAL,
                 "vad: out_stride too small");
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "vad: no CUDA device");
    if (n_clips == 0 || max_frames == 0) return MMLA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VadArgs a;
    a.pcm = pcm;
    a.clip_len_dev = clip_len_dev;
    a.speech = speech;
    a.keep = keep;
    a.voiced_len = voiced_len;
    a.n_clips = n_clips;
    a.clip_stride = clip_stride;
    a.clips_per_stream = clips_per_stream;
    a.n_streams = (n_clips + clips_per_stream - 1) / clips_per_stream;
    a.clip_len = clip_len;
    a.max_frames = max_frames;
    const long long grid = (a.n_streams + 31) / 32;
    MMLA_REQUIRE(grid < (1LL << 31), MMLA_EUNSUP, "vad: too many streams");
    vad_decide_kernel<<<static_cast<unsigned>(grid), 32, 0, st>>>(a);
    mmla_count_launch("vad_decide_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    if (pcm_out) {
        const int vec_ok = ((clip_stride & 7) == 0 && (out_stride & 7) == 0 && (reinterpret_cast<uintptr_t>(pcm) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(pcm_out) & 15) == 0) ? 1 : 0;
        vad_compact_kernel<<<static_cast<unsigned>(n_clips), 128, 0, st>>>(pcm, clip_stride, keep, max_frames, pcm_out, out_stride, vec_ok);
        mmla_count_launch("vad_compact_kernel", st);
        MMLA_CUDA_CHECK(cudaGetLastError());
    }
    return MMLA_OK;
}
