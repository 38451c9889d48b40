// rfi_synth.cu -- device-side synthetic visibilities: the step BEFORE the hot path
// (rfi_toolbox/data_generation/synthetic_generator.py:520-656 `_generate_single_sample`,
// :657-673 `_generate_bandpass`, :675-815 `_add_*`).
//
// Distribution parity only: the reference draws from the host MT19937 stream, which a device
// cannot reproduce; here every pixel owns a Philox4x32-10 counter (key = seed, counter =
// (pixel, baseline, stream)), so a cube is a pure function of (seed, baseline, shape) whatever
// the launch geometry or the baseline sharding over GPUs.
//
// The host draws the (few hundred) RFI events of a baseline exactly as the reference's `_add_*`
// do and hands them over in SEPARABLE form -- the reference's events are all rows x times
// rectangles (plus the sweep): full-row amplitudes, full-column amplitudes, and per narrow band
// a time profile -- so one pass writes every polarisation of a pixel and its mask:
//
//   base      = N(noise, 0.1 noise) * bandpass(row)                          (:549-554)
//   sig       = row_amp[r] + col_amp[t] + sum_k [r in band k] band_amp[k][t] + sweeps(r, t)
//   pol 0     = base + sig ; pol 1 = corr * sig + (1 - corr) * N(0, 0.1 noise) + base ;
//   pol >= 2  = N(noise, 0.1 noise), unflagged                               (:616-636)
//   out       = pol_real * exp(i * U(0, 2 pi)), complex64                    (:639-640)
//
// 9 bytes written per (pixel, pol), nothing of cube size read.  Measured (45 x 4 x 1024 x 1024):
// 0.77 ms, 2.1 TB/s -- issue-bound (77 % issue-active: two Philox4x32-10 blocks, two Box-Muller
// pairs and four sincos per pixel), i.e. 245 Gpix/s, far from the critical path of any bench.
#include "rfi_common.cuh"

namespace rfi {

struct Philox {
    uint32_t k0, k1;
    RFI_DEVINL uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
            const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += W0; b += W1;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

RFI_DEVINL float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)

// two standard normals from two uniforms (Box-Muller)
RFI_DEVINL void normal2(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float r = sqrtf(-2.0f * __logf(u01(a)));
    float s, c;
    __sincosf(6.283185307179586f * u01(b), &s, &c);
    z0 = r * c; z1 = r * s;
}

constexpr int kSynthMaxBands = 64;

struct SynthDev {
    long long C, T;
    int n_pol, n_bands, n_sweeps, bandpass, order, edge;
    float noise, sigma, corr;
    uint32_t k0, k1;
    long long first_baseline;
};

__global__ void __launch_bounds__(256)
synth_kernel(SynthDev s, const float* __restrict__ row_amp, const float* __restrict__ col_amp,
             const int* __restrict__ band_rows, const float* __restrict__ band_amp,
             const float* __restrict__ sweep, float2* __restrict__ cube, uint8_t* __restrict__ mask) {
    // grid: x = column blocks of 256, y = row, z = baseline (local index)
    const long long b = blockIdx.z;
    const int r = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ int act[kSynthMaxBands];   // bands covering this row
    __shared__ int n_act;
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < s.n_bands; ++k) {
            const int r0 = band_rows[(b * s.n_bands + k) * 2], r1 = band_rows[(b * s.n_bands + k) * 2 + 1];
            if (r >= r0 && r < r1) act[n++] = k;
        }
        n_act = n;
    }
    __syncthreads();
    if (t >= s.T) return;

    float bp = 1.0f;
    if (s.bandpass && s.edge > 0) {
        if (r < s.edge) bp = powf((float)r / (float)s.edge, (float)s.order);
        else if (r >= s.C - s.edge) bp = powf((float)(s.C - 1 - r) / (float)s.edge, (float)s.order);
    }
    float sig = row_amp[b * s.C + r] + col_amp[b * s.T + t];
    for (int i = 0; i < n_act; ++i) sig += band_amp[((size_t)b * s.n_bands + act[i]) * s.T + t];
    for (int k = 0; k < s.n_sweeps; ++k) {
        const float* sw = sweep + ((size_t)b * s.n_sweeps + k) * 6;
        float prog = (float)t / (float)s.T;
        if (sw[3] == 2.0f) prog = prog * prog;
        const int centre = (int)(sw[0] + (sw[1] - sw[0]) * prog);   // int() truncation, :786
        const int w2 = (int)sw[2] / 2;
        const int lo = max(0, centre - w2), hi = min((int)s.C, centre + w2);
        if (r >= lo && r < hi) sig += sw[4];
    }
    const bool flagged = sig > 0.0f;

    const Philox ph{s.k0, s.k1};
    const unsigned long long pix = (unsigned long long)r * (unsigned long long)s.T + (unsigned long long)t;
    const uint32_t c0 = (uint32_t)pix, c1 = (uint32_t)(pix >> 32), c2 = (uint32_t)(s.first_baseline + b);
    const uint4 un = ph(c0, c1, c2, 0u), up = ph(c0, c1, c2, 1u);
    float z0, z1, z2, z3;
    normal2(un.x, un.y, z0, z1);
    normal2(un.z, un.w, z2, z3);
    const float base = (s.noise + s.sigma * z0) * bp;
    const size_t plane = (size_t)s.C * s.T;
    const size_t at0 = ((size_t)b * s.n_pol) * plane + (size_t)r * s.T + t;
    auto emit = [&](int p, float real, uint32_t phase_bits) {
        float sn, cs;
        __sincosf(6.283185307179586f * u01(phase_bits), &sn, &cs);
        cube[at0 + p * plane] = make_float2(real * cs, real * sn);
        mask[at0 + p * plane] = (p <= 1 && flagged) ? 1 : 0;
    };
    emit(0, base + sig, up.x);
    if (s.n_pol > 1) emit(1, s.corr * sig + (1.0f - s.corr) * (s.sigma * z1) + base, up.y);
    if (s.n_pol > 2) emit(2, s.noise + s.sigma * z2, up.z);
    if (s.n_pol > 3) emit(3, s.noise + s.sigma * z3, up.w);
    for (int p = 4; p < s.n_pol; ++p) {  // further polarisations: one more counter each
        const uint4 e = ph(c0, c1, c2, (uint32_t)(p - 2));
        float zz, dummy;
        normal2(e.x, e.y, zz, dummy);
        emit(p, s.noise + s.sigma * zz, e.z);
    }
}

}  // namespace rfi

using namespace rfi;

extern "C" int rfi_synth_waterfalls(const rfi_synth_t* sp, int64_t n_baselines, int64_t first_baseline,
                                    const float* row_amp, const float* col_amp, const int32_t* band_rows,
                                    const float* band_amp, const float* sweep, void* cube, uint8_t* mask,
                                    void* stream) {
    if (!sp) { set_error("synth parameters are NULL"); return RFI_E_INVALID; }
    if (sp->channels <= 0 || sp->times <= 0 || sp->n_pol <= 0 || n_baselines < 0) { set_error("bad cube shape"); return RFI_E_INVALID; }
    if (sp->n_bands < 0 || sp->n_bands > kSynthMaxBands) { set_error("at most %d narrow bands per baseline", kSynthMaxBands); return RFI_E_INVALID; }
    if (sp->n_sweeps < 0) { set_error("bad sweep count"); return RFI_E_INVALID; }
    if (n_baselines == 0) return RFI_OK;
    if (!row_amp || !col_amp || !cube || !mask || (sp->n_bands && (!band_rows || !band_amp)) || (sp->n_sweeps && !sweep)) {
        set_error("NULL buffer"); return RFI_E_INVALID; }
    if (sp->channels > 65535 || n_baselines > 65535) { set_error("at most 65535 channels / baselines per call"); return RFI_E_UNSUPPORTED; }
    SynthDev s;
    s.C = sp->channels; s.T = sp->times; s.n_pol = sp->n_pol; s.n_bands = sp->n_bands; s.n_sweeps = sp->n_sweeps;
    s.bandpass = sp->enable_bandpass; s.order = sp->bandpass_order;
    s.edge = (int)((double)sp->channels * 0.1);   // int(num_channels * edge_fraction), :660-661
    s.noise = sp->noise_level; s.sigma = sp->noise_level * 0.1f; s.corr = sp->pol_corr;
    s.k0 = (uint32_t)sp->seed; s.k1 = (uint32_t)(sp->seed >> 32);
    s.first_baseline = first_baseline;
    const dim3 grid((unsigned)((sp->times + 255) / 256), (unsigned)sp->channels, (unsigned)n_baselines);
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(s, row_amp, col_amp, band_rows, band_amp, sweep,
                                                        static_cast<float2*>(cube), mask);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
