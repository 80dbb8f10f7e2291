/*
 * Oracle: the WebRTC voice-activity detector as `webrtcvad.Vad(3).is_speech(frame, 16000)` runs it.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load the library built from this file; the product (mmla_audio_b200/csrc/vad.cu) is a separate device implementation.
 *
 * PARITY UNPINNED.  `webrtcvad` is an un-pinned, un-vendored dependency of the reference (setup.py:32-41; imported at
 * OverlapDetection/scripts/record_on_pc.py:14 and both *_post_processing.py files; instantiated as the module-global
 * `vad = webrtcvad.Vad(3)`, record_on_pc.py:33).  Its only release line (py-webrtcvad 2.0.x) wraps the fixed-point VAD of
 * the WebRTC code base (common_audio/vad/{vad_core,vad_filterbank,vad_gmm,vad_sp}.c and signal_processing helpers).  That
 * code cannot be fetched or built here, so this file RESTATES its published algorithm from memory, function by function
 * under the upstream names, for the one configuration the reference uses:
 *      sample rate 16 kHz, 30 ms frames (480 samples), aggressiveness mode 3,
 * anchored on the reference call sites
 *      frame_generator / vad_collector   OverlapDetection/scripts/record_on_pc.py:229-295
 *      is_speech(frame.bytes, sample_rate)                    record_on_pc.py:254
 * All arithmetic is integer with the C semantics the upstream relies on (int16_t truncation on assignment, arithmetic
 * right shift of negative values, wrap-around int32 products).
 *
 * The detector is STATEFUL: GMM means / stds, minimum trackers, filter memories and the hang-over counter carry from one
 * frame to the next, and — because the reference keeps ONE module-global Vad object — from one clip to the next.
 * `vad_oracle_reset` + repeated `vad_oracle_is_speech` reproduce that; resetting per clip gives the clip-parallel mode.
 */
#include <stdint.h>
#include <string.h>

enum { kNumChannels = 6, kNumGaussians = 2, kTableSize = 12, kMinEnergy = 10 };

typedef struct {
    int32_t downsampling_filter_states[4];
    int16_t noise_means[kTableSize], speech_means[kTableSize], noise_stds[kTableSize], speech_stds[kTableSize];
    int32_t frame_counter;
    int16_t over_hang, num_of_speech;
    int16_t index_vector[16 * kNumChannels], low_value_vector[16 * kNumChannels];
    int16_t mean_value[kNumChannels];
    int16_t upper_state[5], lower_state[5];
    int16_t hp_filter_state[4];
    int16_t over_hang_max_1, over_hang_max_2, individual, total;   /* the 30 ms entries of mode 3 */
} VadInst;

/* ---- vad_core.c tables ---- */
static const int16_t kSpectrumWeight[kNumChannels] = {6, 8, 10, 12, 14, 16};
static const int16_t kNoiseUpdateConst = 655;    /* Q15 */
static const int16_t kSpeechUpdateConst = 6554;  /* Q15 */
static const int16_t kBackEta = 154;             /* Q8  */
static const int16_t kMinimumDifference[kNumChannels] = {544, 544, 576, 576, 576, 576};
static const int16_t kMaximumSpeech[kNumChannels] = {11392, 11392, 11520, 11520, 11520, 11520};
static const int16_t kMinimumMean[kNumGaussians] = {640, 768};
static const int16_t kMaximumNoise[kNumChannels] = {9216, 9088, 8960, 8832, 8704, 8576};
static const int16_t kNoiseDataWeights[kTableSize] = {34, 62, 72, 66, 53, 25, 94, 66, 56, 62, 75, 103};
static const int16_t kSpeechDataWeights[kTableSize] = {48, 82, 45, 87, 50, 47, 80, 46, 83, 41, 78, 81};
static const int16_t kNoiseDataMeans[kTableSize] = {6738, 4892, 7065, 6715, 6771, 3369, 7646, 3863, 7820, 7266, 5020, 4362};
static const int16_t kSpeechDataMeans[kTableSize] = {8306, 10085, 10078, 11823, 11843, 6309, 9473, 9571, 10879, 7581, 8180, 7483};
static const int16_t kNoiseDataStds[kTableSize] = {378, 1064, 493, 582, 688, 593, 474, 697, 475, 688, 421, 455};
static const int16_t kSpeechDataStds[kTableSize] = {555, 505, 567, 524, 585, 1231, 509, 828, 492, 1540, 1079, 850};
static const int16_t kMaxSpeechFrames = 6;
static const int16_t kMinStd = 384;
/* mode 3 ("very aggressive"), 30 ms column: kOverHangMax1VAG[2], kOverHangMax2VAG[2], kLocalThresholdVAG[2], kGlobalThresholdVAG[2] */
enum { kOverHangMax1 = 2, kOverHangMax2 = 3, kLocalThreshold = 94, kGlobalThreshold = 1100 };

/* ---- signal_processing helpers ---- */
static int clz32(uint32_t x) { int n = 0; if (x == 0) return 32; while (!(x & 0x80000000u)) { x <<= 1; ++n; } return n; }
static int16_t SplNormW32(int32_t a) { if (a == 0) return 0; if (a < 0) a = ~a; return (int16_t)(clz32((uint32_t)a) - 1); }
static int16_t SplNormU32(uint32_t a) { if (a == 0) return 0; return (int16_t)clz32(a); }
static int16_t SplGetSizeInBits(uint32_t n) { return (int16_t)(32 - clz32(n)); }
static int32_t SplDivW32W16(int32_t num, int16_t den) { return den != 0 ? (int32_t)(num / den) : (int32_t)0x7FFFFFFF; }
static int32_t MulWrap(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

static int16_t SplGetScalingSquare(const int16_t* v, int n, int times) {
    int16_t nbits = SplGetSizeInBits((uint32_t)times);
    int16_t smax = -1, sabs, t;
    for (int i = 0; i < n; ++i) {
        sabs = (int16_t)(v[i] > 0 ? v[i] : -v[i]);      /* -(-32768) wraps back to -32768, as upstream */
        smax = sabs > smax ? sabs : smax;
    }
    t = SplNormW32((int32_t)smax * smax);
    if (smax == 0) return 0;
    return (int16_t)((t > nbits) ? 0 : nbits - t);
}
static int32_t SplEnergy(const int16_t* v, int n, int* scale_factor) {
    int32_t en = 0;
    int scaling = SplGetScalingSquare(v, n, n);
    for (int i = 0; i < n; ++i) en += ((int32_t)v[i] * v[i]) >> scaling;
    *scale_factor = scaling;
    return en;
}

/* ---- vad_sp.c ---- */
static const int16_t kAllPassCoefsQ13[2] = {5243, 1392};
static const int16_t kSmoothingDown = 6553, kSmoothingUp = 32439;

static void Downsampling(const int16_t* in, int16_t* out, int32_t* filter_state, int in_length) {
    int16_t tmp16_1, tmp16_2;
    int32_t tmp32_1 = filter_state[0], tmp32_2 = filter_state[1];
    int half = in_length >> 1;
    for (int n = 0; n < half; ++n) {
        tmp16_1 = (int16_t)((tmp32_1 >> 1) + ((kAllPassCoefsQ13[0] * *in) >> 14));
        *out = tmp16_1;
        tmp32_1 = (int32_t)(*in++) - ((kAllPassCoefsQ13[0] * tmp16_1) >> 12);
        tmp16_2 = (int16_t)((tmp32_2 >> 1) + ((kAllPassCoefsQ13[1] * *in) >> 14));
        *out = (int16_t)(*out + tmp16_2);
        ++out;
        tmp32_2 = (int32_t)(*in++) - ((kAllPassCoefsQ13[1] * tmp16_2) >> 12);
    }
    filter_state[0] = tmp32_1;
    filter_state[1] = tmp32_2;
}

static int16_t FindMinimum(VadInst* self, int16_t feature_value, int channel) {
    int i, j, position = -1;
    const int offset = channel << 4;
    int16_t current_median = 1600, alpha = 0;
    int32_t tmp32;
    int16_t* age = &self->index_vector[offset];
    int16_t* smallest_values = &self->low_value_vector[offset];
    for (i = 0; i < 16; i++) {
        if (age[i] != 100) {
            age[i]++;
        } else {
            for (j = i; j < 15; j++) {
                smallest_values[j] = smallest_values[j + 1];
                age[j] = age[j + 1];
            }
            age[15] = 101;
            smallest_values[15] = 10000;
        }
    }
    if (feature_value < smallest_values[7]) {
        if (feature_value < smallest_values[3]) {
            if (feature_value < smallest_values[1]) position = feature_value < smallest_values[0] ? 0 : 1;
            else position = feature_value < smallest_values[2] ? 2 : 3;
        } else if (feature_value < smallest_values[5]) {
            position = feature_value < smallest_values[4] ? 4 : 5;
        } else {
            position = feature_value < smallest_values[6] ? 6 : 7;
        }
    } else if (feature_value < smallest_values[15]) {
        if (feature_value < smallest_values[11]) {
            if (feature_value < smallest_values[9]) position = feature_value < smallest_values[8] ? 8 : 9;
            else position = feature_value < smallest_values[10] ? 10 : 11;
        } else if (feature_value < smallest_values[13]) {
            position = feature_value < smallest_values[12] ? 12 : 13;
        } else {
            position = feature_value < smallest_values[14] ? 14 : 15;
        }
    }
    if (position > -1) {
        for (i = 15; i > position; i--) {
            smallest_values[i] = smallest_values[i - 1];
            age[i] = age[i - 1];
        }
        smallest_values[position] = feature_value;
        age[position] = 1;
    }
    if (self->frame_counter > 2) current_median = smallest_values[2];
    else if (self->frame_counter > 0) current_median = smallest_values[0];
    if (self->frame_counter > 0) alpha = current_median < self->mean_value[channel] ? kSmoothingDown : kSmoothingUp;
    tmp32 = (alpha + 1) * self->mean_value[channel];
    tmp32 += (32767 - alpha) * current_median;
    tmp32 += 16384;
    self->mean_value[channel] = (int16_t)(tmp32 >> 15);
    return self->mean_value[channel];
}

/* ---- vad_filterbank.c ---- */
static const int16_t kLogConst = 24660, kLogEnergyIntPart = 14336;
static const int16_t kHpZeroCoefs[3] = {6631, -13262, 6631};
static const int16_t kHpPoleCoefs[3] = {16384, -7756, 5620};
static const int16_t kAllPassCoefsQ15[2] = {20972, 5571};
static const int16_t kOffsetVector[6] = {368, 368, 272, 176, 176, 176};

static void HighPassFilter(const int16_t* in, int n, int16_t* st, int16_t* out) {
    for (int i = 0; i < n; i++) {
        int32_t tmp32 = kHpZeroCoefs[0] * *in;
        tmp32 += kHpZeroCoefs[1] * st[0];
        tmp32 += kHpZeroCoefs[2] * st[1];
        st[1] = st[0];
        st[0] = *in++;
        tmp32 -= kHpPoleCoefs[1] * st[2];
        tmp32 -= kHpPoleCoefs[2] * st[3];
        st[3] = st[2];
        st[2] = (int16_t)(tmp32 >> 14);
        *out++ = st[2];
    }
}
static void AllPassFilter(const int16_t* in, int n, int16_t coef, int16_t* filter_state, int16_t* out) {
    int32_t state32 = (int32_t)(*filter_state) * (1 << 16);
    for (int i = 0; i < n; i++) {
        int32_t tmp32 = state32 + coef * *in;
        int16_t tmp16 = (int16_t)(tmp32 >> 16);
        *out++ = tmp16;
        state32 = (*in * (1 << 14)) - coef * tmp16;
        state32 = (int32_t)((uint32_t)state32 * 2u);
        in += 2;
    }
    *filter_state = (int16_t)(state32 >> 16);
}
static void SplitFilter(const int16_t* in, int n, int16_t* upper_state, int16_t* lower_state, int16_t* hp, int16_t* lp) {
    int half = n >> 1;
    AllPassFilter(&in[0], half, kAllPassCoefsQ15[0], upper_state, hp);
    AllPassFilter(&in[1], half, kAllPassCoefsQ15[1], lower_state, lp);
    for (int i = 0; i < half; i++) {
        int16_t tmp_out = hp[i];
        hp[i] = (int16_t)(hp[i] - lp[i]);
        lp[i] = (int16_t)(lp[i] + tmp_out);
    }
}
static void LogOfEnergy(const int16_t* in, int n, int16_t offset, int16_t* total_energy, int16_t* log_energy) {
    int tot_rshifts = 0;
    uint32_t energy = (uint32_t)SplEnergy(in, n, &tot_rshifts);
    if (energy != 0) {
        int normalizing_rshifts = 17 - SplNormU32(energy);
        int16_t log2_energy = kLogEnergyIntPart;
        tot_rshifts += normalizing_rshifts;
        if (normalizing_rshifts < 0) energy <<= -normalizing_rshifts;
        else energy >>= normalizing_rshifts;
        log2_energy = (int16_t)(log2_energy + (int16_t)((energy & 0x00003FFF) >> 4));
        *log_energy = (int16_t)(((kLogConst * log2_energy) >> 19) + ((tot_rshifts * kLogConst) >> 9));
        if (*log_energy < 0) *log_energy = 0;
    } else {
        *log_energy = offset;
        return;
    }
    *log_energy = (int16_t)(*log_energy + offset);
    if (*total_energy <= kMinEnergy) {
        if (tot_rshifts >= 0) *total_energy = (int16_t)(*total_energy + kMinEnergy + 1);
        else *total_energy = (int16_t)(*total_energy + (int16_t)(energy >> -tot_rshifts));
    }
}
static int16_t CalculateFeatures(VadInst* self, const int16_t* data_in, int data_length, int16_t* features) {
    int16_t total_energy = 0;
    int16_t hp_120[120], lp_120[120], hp_60[60], lp_60[60];
    const int half_data_length = data_length >> 1;
    int length = half_data_length;
    SplitFilter(data_in, data_length, &self->upper_state[0], &self->lower_state[0], hp_120, lp_120);   /* split at 2000 Hz */
    SplitFilter(hp_120, length, &self->upper_state[1], &self->lower_state[1], hp_60, lp_60);           /* 2-4 kHz at 3000 */
    length >>= 1;
    LogOfEnergy(hp_60, length, kOffsetVector[5], &total_energy, &features[5]);                           /* 3000-4000 */
    LogOfEnergy(lp_60, length, kOffsetVector[4], &total_energy, &features[4]);                           /* 2000-3000 */
    length = half_data_length;
    SplitFilter(lp_120, length, &self->upper_state[2], &self->lower_state[2], hp_60, lp_60);           /* 0-2 kHz at 1000 */
    length >>= 1;
    LogOfEnergy(hp_60, length, kOffsetVector[3], &total_energy, &features[3]);                           /* 1000-2000 */
    SplitFilter(lp_60, length, &self->upper_state[3], &self->lower_state[3], hp_120, lp_120);          /* 0-1 kHz at 500 */
    length >>= 1;
    LogOfEnergy(hp_120, length, kOffsetVector[2], &total_energy, &features[2]);                          /* 500-1000 */
    SplitFilter(lp_120, length, &self->upper_state[4], &self->lower_state[4], hp_60, lp_60);           /* 0-500 at 250 */
    length >>= 1;
    LogOfEnergy(hp_60, length, kOffsetVector[1], &total_energy, &features[1]);                           /* 250-500 */
    HighPassFilter(lp_60, length, self->hp_filter_state, hp_120);                                        /* remove 0-80 */
    LogOfEnergy(hp_120, length, kOffsetVector[0], &total_energy, &features[0]);                          /* 80-250 */
    return total_energy;
}

/* ---- vad_gmm.c ---- */
static const int32_t kCompVar = 22005;
static const int16_t kLog2Exp = 5909;
static int32_t GaussianProbability(int16_t input, int16_t mean, int16_t std, int16_t* delta) {
    int16_t tmp16, inv_std, inv_std2, exp_value = 0;
    int32_t tmp32;
    tmp32 = (int32_t)131072 + (int32_t)(std >> 1);
    inv_std = (int16_t)SplDivW32W16(tmp32, std);
    tmp16 = (int16_t)(inv_std >> 2);
    inv_std2 = (int16_t)((tmp16 * tmp16) >> 2);
    tmp16 = (int16_t)(input << 3);
    tmp16 = (int16_t)(tmp16 - mean);
    *delta = (int16_t)((inv_std2 * tmp16) >> 10);
    tmp32 = (*delta * tmp16) >> 9;
    if (tmp32 < kCompVar) {
        tmp16 = (int16_t)((kLog2Exp * tmp32) >> 12);
        tmp16 = (int16_t)(-tmp16);
        exp_value = (int16_t)(0x0400 | (tmp16 & 0x03FF));
        tmp16 = (int16_t)(tmp16 ^ 0xFFFF);
        tmp16 >>= 10;
        tmp16 = (int16_t)(tmp16 + 1);
        exp_value >>= tmp16;
    }
    return inv_std * exp_value;
}

/* ---- vad_core.c ---- */
static int32_t WeightedAverage(int16_t* data, int16_t offset, const int16_t* weights) {
    int32_t weighted_average = 0;
    for (int k = 0; k < kNumGaussians; k++) {
        data[k * kNumChannels] = (int16_t)(data[k * kNumChannels] + offset);
        weighted_average += data[k * kNumChannels] * weights[k * kNumChannels];
    }
    return weighted_average;
}

static int16_t GmmProbability(VadInst* self, int16_t* features, int16_t total_power) {
    int channel, k, gaussian;
    int16_t feature_minimum, h0, h1, log_likelihood_ratio, vadflag = 0, shifts_h0, shifts_h1;
    int16_t tmp_s16, tmp1_s16, tmp2_s16, diff, nmk, nmk2, nmk3, smk, smk2, nsk, ssk, delt, ndelt, maxspe, maxmu;
    int16_t deltaN[kTableSize], deltaS[kTableSize];
    int16_t ngprvec[kTableSize] = {0}, sgprvec[kTableSize] = {0};
    int32_t h0_test, h1_test, tmp1_s32, tmp2_s32, sum_log_likelihood_ratios = 0, noise_global_mean, speech_global_mean;
    int32_t noise_probability[kNumGaussians], speech_probability[kNumGaussians];
    const int16_t overhead1 = self->over_hang_max_1, overhead2 = self->over_hang_max_2;
    const int16_t individualTest = self->individual, totalTest = self->total;

    if (total_power > kMinEnergy) {
        for (channel = 0; channel < kNumChannels; channel++) {
            h0_test = 0;
            h1_test = 0;
            for (k = 0; k < kNumGaussians; k++) {
                gaussian = channel + k * kNumChannels;
                tmp1_s32 = GaussianProbability(features[channel], self->noise_means[gaussian], self->noise_stds[gaussian], &deltaN[gaussian]);
                noise_probability[k] = kNoiseDataWeights[gaussian] * tmp1_s32;
                h0_test += noise_probability[k];
                tmp1_s32 = GaussianProbability(features[channel], self->speech_means[gaussian], self->speech_stds[gaussian], &deltaS[gaussian]);
                speech_probability[k] = kSpeechDataWeights[gaussian] * tmp1_s32;
                h1_test += speech_probability[k];
            }
            shifts_h0 = SplNormW32(h0_test);
            shifts_h1 = SplNormW32(h1_test);
            if (h0_test == 0) shifts_h0 = 31;
            if (h1_test == 0) shifts_h1 = 31;
            log_likelihood_ratio = (int16_t)(shifts_h0 - shifts_h1);
            sum_log_likelihood_ratios += (int32_t)(log_likelihood_ratio * kSpectrumWeight[channel]);
            if ((log_likelihood_ratio * 4) > individualTest) vadflag = 1;
            h0 = (int16_t)(h0_test >> 12);
            if (h0 > 0) {
                tmp1_s32 = (int32_t)(((uint32_t)noise_probability[0] & 0xFFFFF000u) << 2);
                ngprvec[channel] = (int16_t)SplDivW32W16(tmp1_s32, h0);
                ngprvec[channel + kNumChannels] = (int16_t)(16384 - ngprvec[channel]);
            } else {
                ngprvec[channel] = 16384;
            }
            h1 = (int16_t)(h1_test >> 12);
            if (h1 > 0) {
                tmp1_s32 = (int32_t)(((uint32_t)speech_probability[0] & 0xFFFFF000u) << 2);
                sgprvec[channel] = (int16_t)SplDivW32W16(tmp1_s32, h1);
                sgprvec[channel + kNumChannels] = (int16_t)(16384 - sgprvec[channel]);
            }
        }
        vadflag |= (sum_log_likelihood_ratios >= totalTest);

        maxspe = 12800;
        for (channel = 0; channel < kNumChannels; channel++) {
            feature_minimum = FindMinimum(self, features[channel], channel);
            noise_global_mean = WeightedAverage(&self->noise_means[channel], 0, &kNoiseDataWeights[channel]);
            tmp1_s16 = (int16_t)(noise_global_mean >> 6);
            for (k = 0; k < kNumGaussians; k++) {
                gaussian = channel + k * kNumChannels;
                nmk = self->noise_means[gaussian];
                smk = self->speech_means[gaussian];
                nsk = self->noise_stds[gaussian];
                ssk = self->speech_stds[gaussian];
                nmk2 = nmk;
                if (!vadflag) {
                    delt = (int16_t)((ngprvec[gaussian] * deltaN[gaussian]) >> 11);
                    nmk2 = (int16_t)(nmk + (int16_t)((delt * kNoiseUpdateConst) >> 22));
                }
                ndelt = (int16_t)((feature_minimum << 4) - tmp1_s16);
                nmk3 = (int16_t)(nmk2 + (int16_t)((ndelt * kBackEta) >> 9));
                tmp_s16 = (int16_t)((k + 5) << 7);
                if (nmk3 < tmp_s16) nmk3 = tmp_s16;
                tmp_s16 = (int16_t)((72 + k - channel) << 7);
                if (nmk3 > tmp_s16) nmk3 = tmp_s16;
                self->noise_means[gaussian] = nmk3;
                if (vadflag) {
                    delt = (int16_t)((sgprvec[gaussian] * deltaS[gaussian]) >> 11);
                    tmp_s16 = (int16_t)((delt * kSpeechUpdateConst) >> 21);
                    smk2 = (int16_t)(smk + ((tmp_s16 + 1) >> 1));
                    maxmu = (int16_t)(maxspe + 640);
                    if (smk2 < kMinimumMean[k]) smk2 = kMinimumMean[k];
                    if (smk2 > maxmu) smk2 = maxmu;
                    self->speech_means[gaussian] = smk2;
                    tmp_s16 = (int16_t)((smk + 4) >> 3);
                    tmp_s16 = (int16_t)(features[channel] - tmp_s16);
                    tmp1_s32 = (deltaS[gaussian] * tmp_s16) >> 3;
                    tmp2_s32 = tmp1_s32 - 4096;
                    tmp_s16 = (int16_t)(sgprvec[gaussian] >> 2);
                    tmp1_s32 = MulWrap(tmp_s16, tmp2_s32);
                    tmp2_s32 = tmp1_s32 >> 4;
                    if (tmp2_s32 > 0) {
                        tmp_s16 = (int16_t)SplDivW32W16(tmp2_s32, (int16_t)(ssk * 10));
                    } else {
                        tmp_s16 = (int16_t)SplDivW32W16(-tmp2_s32, (int16_t)(ssk * 10));
                        tmp_s16 = (int16_t)(-tmp_s16);
                    }
                    tmp_s16 = (int16_t)(tmp_s16 + 128);
                    ssk = (int16_t)(ssk + (tmp_s16 >> 8));
                    if (ssk < kMinStd) ssk = kMinStd;
                    self->speech_stds[gaussian] = ssk;
                } else {
                    tmp_s16 = (int16_t)(features[channel] - (nmk >> 3));
                    tmp1_s32 = (deltaN[gaussian] * tmp_s16) >> 3;
                    tmp1_s32 -= 4096;
                    tmp_s16 = (int16_t)((ngprvec[gaussian] + 2) >> 2);
                    tmp2_s32 = MulWrap(tmp_s16, tmp1_s32);
                    tmp1_s32 = tmp2_s32 >> 14;
                    if (tmp1_s32 > 0) {
                        tmp_s16 = (int16_t)SplDivW32W16(tmp1_s32, nsk);
                    } else {
                        tmp_s16 = (int16_t)SplDivW32W16(-tmp1_s32, nsk);
                        tmp_s16 = (int16_t)(-tmp_s16);
                    }
                    tmp_s16 = (int16_t)(tmp_s16 + 32);
                    nsk = (int16_t)(nsk + (tmp_s16 >> 6));
                    if (nsk < kMinStd) nsk = kMinStd;
                    self->noise_stds[gaussian] = nsk;
                }
            }
            noise_global_mean = WeightedAverage(&self->noise_means[channel], 0, &kNoiseDataWeights[channel]);
            speech_global_mean = WeightedAverage(&self->speech_means[channel], 0, &kSpeechDataWeights[channel]);
            diff = (int16_t)((int16_t)(speech_global_mean >> 9) - (int16_t)(noise_global_mean >> 9));
            if (diff < kMinimumDifference[channel]) {
                tmp_s16 = (int16_t)(kMinimumDifference[channel] - diff);
                tmp1_s16 = (int16_t)((13 * tmp_s16) >> 2);
                tmp2_s16 = (int16_t)((3 * tmp_s16) >> 2);
                speech_global_mean = WeightedAverage(&self->speech_means[channel], tmp1_s16, &kSpeechDataWeights[channel]);
                noise_global_mean = WeightedAverage(&self->noise_means[channel], (int16_t)(-tmp2_s16), &kNoiseDataWeights[channel]);
            }
            maxspe = kMaximumSpeech[channel];
            tmp2_s16 = (int16_t)(speech_global_mean >> 7);
            if (tmp2_s16 > maxspe) {
                tmp2_s16 = (int16_t)(tmp2_s16 - maxspe);
                for (k = 0; k < kNumGaussians; k++) self->speech_means[channel + k * kNumChannels] = (int16_t)(self->speech_means[channel + k * kNumChannels] - tmp2_s16);
            }
            tmp2_s16 = (int16_t)(noise_global_mean >> 7);
            if (tmp2_s16 > kMaximumNoise[channel]) {
                tmp2_s16 = (int16_t)(tmp2_s16 - kMaximumNoise[channel]);
                for (k = 0; k < kNumGaussians; k++) self->noise_means[channel + k * kNumChannels] = (int16_t)(self->noise_means[channel + k * kNumChannels] - tmp2_s16);
            }
        }
        self->frame_counter++;
    }
    if (!vadflag) {
        if (self->over_hang > 0) {
            vadflag = (int16_t)(2 + self->over_hang);
            self->over_hang--;
        }
        self->num_of_speech = 0;
    } else {
        self->num_of_speech++;
        if (self->num_of_speech > kMaxSpeechFrames) {
            self->num_of_speech = kMaxSpeechFrames;
            self->over_hang = overhead2;
        } else {
            self->over_hang = overhead1;
        }
    }
    return vadflag;
}

/* ---- exported test interface ---- */
int vad_oracle_state_bytes(void) { return (int)sizeof(VadInst); }

/* WebRtcVad_InitCore + WebRtcVad_set_mode_core(3) */
void vad_oracle_reset(void* state) {
    VadInst* self = (VadInst*)state;
    memset(self, 0, sizeof(*self));
    for (int i = 0; i < kTableSize; i++) {
        self->noise_means[i] = kNoiseDataMeans[i];
        self->speech_means[i] = kSpeechDataMeans[i];
        self->noise_stds[i] = kNoiseDataStds[i];
        self->speech_stds[i] = kSpeechDataStds[i];
    }
    for (int i = 0; i < 16 * kNumChannels; i++) {
        self->low_value_vector[i] = 10000;
        self->index_vector[i] = 0;
    }
    for (int i = 0; i < kNumChannels; i++) self->mean_value[i] = 1600;
    self->over_hang_max_1 = kOverHangMax1;
    self->over_hang_max_2 = kOverHangMax2;
    self->individual = kLocalThreshold;
    self->total = kGlobalThreshold;
}

/* webrtcvad.Vad.is_speech(frame, 16000) for one 30 ms frame (480 samples): WebRtcVad_Process -> CalcVad16khz -> CalcVad8khz.
 * `features_out` (may be NULL) receives the six sub-band log energies and the total power for diagnostics. */
int vad_oracle_is_speech(void* state, const int16_t* frame480, int16_t* features_out) {
    VadInst* self = (VadInst*)state;
    int16_t speechNB[240], feature_vector[kNumChannels], total_power;
    Downsampling(frame480, speechNB, &self->downsampling_filter_states[2], 480);
    total_power = CalculateFeatures(self, speechNB, 240, feature_vector);
    if (features_out) {
        for (int i = 0; i < kNumChannels; i++) features_out[i] = feature_vector[i];
        features_out[kNumChannels] = total_power;
    }
    return GmmProbability(self, feature_vector, total_power) > 0 ? 1 : 0;
}

/* All complete 30 ms frames frame_generator() yields for a clip of n samples (record_on_pc.py:229-244:
 * `while offset + n < len(audio)`, byte offsets, so a clip that is an exact multiple of 480 samples loses its last frame). */
int vad_oracle_num_frames(int n_samples) {
    const long nbytes = 2L * n_samples;
    int frames = 0;
    for (long off = 0; off + 960 < nbytes; off += 960) ++frames;
    return frames;
}

int vad_oracle_clip(void* state, const int16_t* pcm, int n_samples, uint8_t* is_speech_out) {
    const int nf = vad_oracle_num_frames(n_samples);
    for (int f = 0; f < nf; ++f) is_speech_out[f] = (uint8_t)vad_oracle_is_speech(state, pcm + 480 * f, 0);
    return nf;
}
