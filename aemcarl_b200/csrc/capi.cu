// capi.cu -- C ABI of libaemcarl_b200.so (declared in include/aemcarl_b200.h).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "aem_common.cuh"

static thread_local char g_err[512] = "";

int aem_set_error(int code, const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int aem_set_cuda_error(cudaError_t e, const char *what, const char *file, int line)
{
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return AEM_ERR_CUDA;
}

namespace {

using namespace aem;

// pack nn.Linear weights W[out][in] into packed[k][g][c] (see aem_common.cuh)
void pack_plain(std::vector<float> &dst, const float *W, int N, int K, int TNH)
{
    const int TNHP = pad4(TNH), NP = NGROUPS * TNHP;
    const size_t base = dst.size();
    dst.resize(base + (size_t)K * NP, 0.0f);
    for (int k = 0; k < K; ++k)
        for (int g = 0; g < NGROUPS; ++g)
            for (int c = 0; c < TNH; ++c) {
                const int col = g * TNH + c;
                if (col < N) dst[base + (size_t)k * NP + g * TNHP + c] = W[(size_t)col * K + k];
            }
}

void pack_two(std::vector<float> &dst, const float *W0, const float *W1, int N, int K, int TNH)
{
    const int TNHP = pad4(TNH), NP = NGROUPS * TNHP;
    const size_t base = dst.size();
    dst.resize(base + (size_t)K * NP, 0.0f);
    for (int k = 0; k < K; ++k)
        for (int g = 0; g < NGROUPS; ++g)
            for (int c = 0; c < TNH; ++c) {
                const int col = (g & 7) * TNH + c;
                const float *W = (g >> 3) ? W1 : W0;
                if (col < N) dst[base + (size_t)k * NP + g * TNHP + c] = W[(size_t)col * K + k];
            }
}


// ---------------------------------------------------------------------------------------------------------
// tensor-core images (lookahead_tc.cu).  For every GEMM: W_perm[c][k] = W[out_map(c)][in_map(k)] (0 where a map
// returns -1), split into hi = round-to-nearest TF32 and lo = w - hi, and laid out as img[k/4][c][k%4].
// Thread `half` (0/1) of a token row owns one contiguous block of the permuted columns.
// ---------------------------------------------------------------------------------------------------------
float tf32_rna_host(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// output maps (c -> real output index)
int out50(int c) { int h = c / 32, j = c % 32; return j < 25 ? 25 * h + j : -1; }            // N_pad 64
int out100(int c) { int h = c / 56, j = c % 56; return j < 50 ? 50 * h + j : -1; }           // N_pad 112
int unit75(int h, int j) { return h == 0 ? (j < 37 ? j : -1) : (j < 38 ? 37 + j : -1); }      // 75 hidden units -> 40 | 40
// input maps (k -> real input index)
int in_hx(int k)    // A1 / A3 layout [h0(25) x0..6 | h1(25) x7..12 0] -> index into [h(50) ; x(13)]
{
    int h = k / 32, j = k % 32;
    if (j < 25) return 25 * h + j;
    int xi = 7 * h + (j - 25);
    return (xi < 13 && !(h == 1 && j == 31)) ? 50 + xi : -1;
}
int in100(int k) { int h = k / 52, j = k % 52; return j < 50 ? 50 * h + j : -1; }            // K_pad 104
int in50(int k) { int h = k / 28, j = k % 28; return j < 25 ? 25 * h + j : -1; }             // K_pad 56
int in_joint(int k) { int h = k / 28, j = k % 28; return j < 25 ? 6 + 25 * h + j : 3 * h + (j - 25); }

struct TcPack {
    std::vector<float> img, pblk;
    aem::TcLayer layers[aem::TL_NUM];
    template <class OM, class IM>
    void add(int id, const float *W, int in_dim, const float *bias, int N, int K, OM om, IM im)
    {
        aem::TcLayer &L = layers[id];
        L.N = N; L.ksteps = K / 8; L.bytes = N * K * 4;
        L.off_hi = (int)img.size();
        img.resize(img.size() + (size_t)2 * N * K, 0.0f);
        L.off_lo = L.off_hi + N * K;
        for (int c = 0; c < N; ++c)
            for (int k = 0; k < K; ++k) {
                const int o = om(c), i = im(k);
                const float w = (o >= 0 && i >= 0) ? W[(size_t)o * in_dim + i] : 0.0f;
                const float hi = tf32_rna_host(w);
                const size_t at = ((size_t)(k / 4) * N + c) * 4 + (k % 4);
                img[L.off_hi + at] = hi;
                img[L.off_lo + at] = w - hi;
            }
        L.bias_off = (int)pblk.size();
        for (int c = 0; c < N; ++c) {
            const int o = om(c);
            pblk.push_back((o >= 0 && bias) ? bias[o] : 0.0f);
        }
    }
};

void pack_tc(const float *w, TcPack &P)
{
    using namespace aem;
    auto step1 = [](int k) { return in_hx(k < 8 ? 24 + k : 56 + (k - 8)); };   // A1 columns {24..31, 56..63}
    P.add(TL_Z1S, w + OFF_Z1W, HXD, w + OFF_Z1B, 112, 16, out100, step1);
    P.add(TL_QS, w + OFF_QW, HXD, w + OFF_QB, 64, 16, out50, step1);
    P.add(TL_Z1, w + OFF_Z1W, HXD, w + OFF_Z1B, 112, 64, out100, in_hx);
    P.add(TL_Z2, w + OFF_Z2W, 100, w + OFF_Z2B, 64, 104, out50, in100);
    P.add(TL_R1, w + OFF_R1W, HXD, w + OFF_R1B, 112, 64, out100, in_hx);
    P.add(TL_R2, w + OFF_R2W, 100, w + OFF_R2B, 64, 104, out50, in100);
    P.add(TL_Q, w + OFF_QW, HXD, w + OFF_QB, 64, 64, out50, in_hx);
    for (int l = 0; l < 3; ++l) {
        const float *E = w + OFF_ENC + l * ENC_SIZE;
        const int LB = TL_ENC + 7 * l;
        for (int hd = 0; hd < 2; ++hd)       // in_proj per head: columns [q(25+7) | k(25+7) | v(25+7)]
            P.add(LB + hd, E + ENC_INW, HID, E + ENC_INB, 96, 56,
                  [hd](int c) { int s = c / 32, j = c % 32; return j < 25 ? s * 50 + 25 * hd + j : -1; }, in50);
        P.add(LB + 2, E + ENC_OW, HID, E + ENC_OB, 64, 56, out50, in50);
        for (int ch = 0; ch < 2; ++ch)       // linear1 in two chunks of 75 hidden units
            P.add(LB + 3 + ch, E + ENC_L1W, HID, E + ENC_L1B, 80, 56,
                  [ch](int c) { int u = unit75(c / 40, c % 40); return u >= 0 ? 75 * ch + u : -1; }, in50);
        for (int ch = 0; ch < 2; ++ch)       // linear2 accumulated over the two chunks; bias applied once (chunk b)
            P.add(LB + 5 + ch, E + ENC_L2W, 150, ch == 1 ? E + ENC_L2B : nullptr, 64, 80, out50,
                  [ch](int k) { int u = unit75(k / 40, k % 40); return u >= 0 ? 75 * ch + u : -1; });
    }
    P.add(TL_A0, w + OFF_A0W, 56, w + OFF_A0B, 112, 56, out100, in_joint);
    P.add(TL_A2, w + OFF_A2W, 100, w + OFF_A2B, 64, 104, out50, in100);
}

// ---------------------------------------------------------------------------------------------------------
// FP16x3 images (lookahead_h16.cu): W_perm[c][k] split into hi = fp16(w), lo = fp16(w - hi), img[k/8][c][k%8].
// N / K layouts: N50 -> [32 | 32], N100 -> [64 | 64] (converted in place to K = 128), linear1 chunks [48 | 48]
// (in place, K = 96), in_proj per head [q32 | k32 | v32]; K = 64 operands are [32 | 32].
// ---------------------------------------------------------------------------------------------------------
int out100w(int c) { int h = c / 64, j = c % 64; return j < 50 ? 50 * h + j : -1; }
int in100w(int k) { int h = k / 64, j = k % 64; return j < 50 ? 50 * h + j : -1; }
int in50w(int k) { int h = k / 32, j = k % 32; return j < 25 ? 25 * h + j : -1; }
int in_jointw(int k) { int h = k / 32, j = k % 32; return j < 25 ? 6 + 25 * h + j : (j < 28 ? 3 * h + (j - 25) : -1); }

struct H16Pack {
    std::vector<__half> img;
    std::vector<float> pblk;
    aem::TcLayer layers[aem::TL_NUM];
    // The bias rides in the weights: K slot `bias_k` of every operand layout is a padding slot that the kernel keeps at
    // exactly 1.0, so column `bias_k` of the image holds the bias (hi/lo split like every weight).  Layers whose
    // accumulators are converted in place into the next operand get a padding neuron `one_c` whose only non-zero weight is
    // 1.0 at `bias_k`: it evaluates to relu(1) = 1 and becomes the next layer's constant slot.
    template <class OM, class IM>
    void add(int id, const float *W, int in_dim, const float *bias, int N, int K, OM om, IM im, int bias_k, int one_c)
    {
        aem::TcLayer &L = layers[id];
        L.N = N; L.ksteps = K / 16; L.bytes = N * K * 2;
        L.off_hi = (int)img.size();
        img.resize(img.size() + (size_t)2 * N * K, __float2half(0.0f));
        L.off_lo = L.off_hi + N * K;
        if (im(bias_k) >= 0 || (one_c >= 0 && om(one_c) >= 0)) { fprintf(stderr, "pack_h16: slot clash at layer %d\n", id); abort(); }
        for (int c = 0; c < N; ++c)
            for (int k = 0; k < K; ++k) {
                const int o = om(c), i = im(k);
                float w = (o >= 0 && i >= 0) ? W[(size_t)o * in_dim + i] : 0.0f;
                if (k == bias_k) w = (o >= 0 && bias) ? bias[o] : (c == one_c ? 1.0f : 0.0f);
                const __half hi = __float2half_rn(w);
                const __half lo = __float2half_rn(w - __half2float(hi));
                const size_t at = ((size_t)(k / 8) * N + c) * 8 + (k % 8);
                img[L.off_hi + at] = hi;
                img[L.off_lo + at] = lo;
            }
        L.bias_off = (N * K * 2 > 12288) ? 2 : 1;     // h16 tables: k-halves per image (ring slots are 12 KB)
    }
};

void pack_h16(const float *w, H16Pack &P)
{
    using namespace aem;
    auto step1 = [](int k) { return in_hx(k < 16 ? 16 + k : 48 + (k - 16)); };     // A1 slots {16..31, 48..63}
    P.add(TL_Z1, w + OFF_Z1W, HXD, w + OFF_Z1B, 128, 64, out100w, in_hx, 63, 127);
    P.add(TL_Z2, w + OFF_Z2W, 100, w + OFF_Z2B, 64, 128, out50, in100w, 127, -1);
    P.add(TL_R1, w + OFF_R1W, HXD, w + OFF_R1B, 128, 64, out100w, in_hx, 63, 127);
    P.add(TL_R2, w + OFF_R2W, 100, w + OFF_R2B, 64, 128, out50, in100w, 127, -1);
    P.add(TL_Q, w + OFF_QW, HXD, w + OFF_QB, 64, 64, out50, in_hx, 63, -1);
    P.add(TL_Z1S, w + OFF_Z1W, HXD, w + OFF_Z1B, 128, 32, out100w, step1, 31, 127);
    P.add(TL_QS, w + OFF_QW, HXD, w + OFF_QB, 64, 32, out50, step1, 31, -1);
    for (int l = 0; l < 3; ++l) {
        const float *E = w + OFF_ENC + l * ENC_SIZE;
        const int LB = TL_ENC + 7 * l;
        for (int hd = 0; hd < 2; ++hd)
            P.add(LB + hd, E + ENC_INW, HID, E + ENC_INB, 96, 64,
                  [hd](int c) { int sidx = c / 32, j = c % 32; return j < 25 ? sidx * 50 + 25 * hd + j : -1; }, in50w, 63, -1);
        P.add(LB + 2, E + ENC_OW, HID, E + ENC_OB, 64, 64, out50, in50w, 63, -1);
        for (int ch = 0; ch < 2; ++ch)
            P.add(LB + 3 + ch, E + ENC_L1W, HID, E + ENC_L1B, 96, 64,
                  [ch](int c) { int u = unit75(c / 48, c % 48); return u >= 0 ? 75 * ch + u : -1; }, in50w, 63, 95);
        for (int ch = 0; ch < 2; ++ch)       // linear2 accumulates the two chunks; the bias rides in chunk b
            P.add(LB + 5 + ch, E + ENC_L2W, 150, ch == 1 ? E + ENC_L2B : nullptr, 64, 96, out50,
                  [ch](int k) { int u = unit75(k / 48, k % 48); return u >= 0 ? 75 * ch + u : -1; }, 95, -1);
    }
    P.add(TL_A0, w + OFF_A0W, 56, w + OFF_A0B, 128, 64, out100w, in_jointw, 63, 127);
    P.add(TL_A2, w + OFF_A2W, 100, w + OFF_A2B, 64, 128, out50, in100w, 127, -1);
    // parameter block (lookahead_h16.cu PB_*): every 50-vector as two rows of 28 floats (25 values + 3 zeros), one per
    // thread half, so that a thread reads its values as 7 float4.
    //   LayerNorm [layer][norm1 / norm2][half][gamma 28 | beta 28]  (672 floats)
    //   ponder    [half][28] + bias + 3 zeros                       (at 672, bias at 728)
    //   action_mlp.4 [half][28] + bias                              (at 732, bias at 788)
    auto rows28 = [&P](const float *v) {
        for (int hf = 0; hf < 2; ++hf)
            for (int j = 0; j < 28; ++j) P.pblk.push_back(j < 25 ? v[25 * hf + j] : 0.0f);
    };
    for (int l = 0; l < 3; ++l) {
        const float *E = w + OFF_ENC + l * ENC_SIZE;
        for (int nm = 0; nm < 2; ++nm) {
            const float *g = E + (nm ? ENC_N2G : ENC_N1G), *b = E + (nm ? ENC_N2B : ENC_N1B);
            for (int hf = 0; hf < 2; ++hf) {
                for (int j = 0; j < 28; ++j) P.pblk.push_back(j < 25 ? g[25 * hf + j] : 0.0f);
                for (int j = 0; j < 28; ++j) P.pblk.push_back(j < 25 ? b[25 * hf + j] : 0.0f);
            }
        }
    }
    rows28(w + OFF_PW);
    P.pblk.push_back(w[OFF_PB]);
    while (P.pblk.size() < 732) P.pblk.push_back(0.0f);
    rows28(w + OFF_A4W);
    P.pblk.push_back(w[OFF_A4B]);
    while (P.pblk.size() < (size_t)H16_PBLK_FLOATS) P.pblk.push_back(0.0f);
}

cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int check_ready(aem_handle h, bool need_weights, bool need_env)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (need_weights && !h->weights_set) return aem_set_error(AEM_ERR_STATE, "aem_load_weights has not been called");
    if (!h->actions_set) return aem_set_error(AEM_ERR_STATE, "aem_set_action_space has not been called");
    if (need_env && !h->params_set) return aem_set_error(AEM_ERR_STATE, "aem_set_env_params has not been called");
    return AEM_OK;
}

double action_amax(aem_handle h)
{
    double amax = 0.0;
    for (int a = 0; a < h->A; ++a) {
        const double s = hypot(h->actions_h[2 * a], h->actions_h[2 * a + 1]) * h->params.agent_timestep;
        if (s > amax) amax = s;
    }
    return amax;
}

}  // namespace

extern "C" {

int aem_version(void) { return AEM_VERSION; }
const char *aem_last_error(void) { return g_err; }

int aem_create(int device, aem_handle *out)
{
    if (!out) return aem_set_error(AEM_ERR_INVALID, "out is null");
    int count = 0;
    AEM_CUDA_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return aem_set_error(AEM_ERR_INVALID, "no such CUDA device");
    AEM_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    AEM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return aem_set_error(AEM_ERR_INVALID, "libaemcarl_b200 is built for sm_100a (B200) only");
    aem_handle h = new aem_handle_s();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    AEM_CUDA_CHECK(cudaMalloc(&h->raw_d, sizeof(float) * NUM_WEIGHTS));
    AEM_CUDA_CHECK(cudaMalloc(&h->actions_d, sizeof(double) * AEM_MAX_ACTIONS * 2));
    AEM_CUDA_CHECK(cudaMalloc(&h->lut_d, sizeof(double) * 10000));
    {   // crowd_sim.py:20-28 norm.pdf((i - 5000) / 1000), same expression as oracle/aem_oracle.c
        std::vector<double> lut(10000);
        for (int i = 0; i < 10000; ++i) {
            const double x = (double)(i - 5000) / 1000.0;
            lut[i] = exp(-x * x / 2.0) / sqrt(2.0 * 3.141592653589793);
        }
        AEM_CUDA_CHECK(cudaMemcpy(h->lut_d, lut.data(), sizeof(double) * 10000, cudaMemcpyHostToDevice));
    }
    h->net.act_fixed = 0;
    h->net.act_steps = 1;
    *out = h;
    return AEM_OK;
}

int aem_destroy(aem_handle h)
{
    if (!h) return AEM_OK;
    cudaSetDevice(h->device);
    cudaFree(h->raw_d); cudaFree(h->packed_d); cudaFree(h->tc_img_d); cudaFree(h->tc_pblk_d); cudaFree(h->tc_layers_d); cudaFree(h->h16_img_d); cudaFree(h->h16_pblk_d); cudaFree(h->h16_layers_d); cudaFree(h->actions_d); cudaFree(h->lut_d);
    cudaFree(h->s_robot); cudaFree(h->s_humans); cudaFree(h->s_time); cudaFree(h->s_newvel);
    cudaFree(h->s_rewards); cudaFree(h->s_dmin); cudaFree(h->s_hnext); cudaFree(h->s_values);
    cudaFree(h->s_oreward); cudaFree(h->s_info); cudaFree(h->s_best); cudaFree(h->s_odone); cudaFree(h->s_oinfo);
    if (h->s_stream) cudaStreamDestroy(as_stream(h->s_stream));
    cudaFree(h->tail_d);
    cudaFree(h->tr_part); cudaFree(h->tr_sqerr); cudaFree(h->tr_stash); cudaFree(h->tr_flags); cudaFree(h->tr_gru);
    delete h;
    return AEM_OK;
}

int aem_load_weights(aem_handle h, const float *w, size_t count)
{
    if (!h || !w) return aem_set_error(AEM_ERR_INVALID, "null argument");
    if (count != (size_t)NUM_WEIGHTS) return aem_set_error(AEM_ERR_INVALID, "weight blob must hold 113752 floats");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    std::vector<float> pk;
    NetParams &net = h->net;
    net.off_z1r1 = (int)pk.size(); pack_two(pk, w + OFF_Z1W, w + OFF_R1W, 100, HXD, TNH_Z1R1);
    net.off_z2r2 = (int)pk.size(); pack_two(pk, w + OFF_Z2W, w + OFF_R2W, HID, 100, TNH_Z2R2);
    net.off_q = (int)pk.size();    pack_plain(pk, w + OFF_QW, HID, HXD, TNH_50);
    for (int l = 0; l < 3; ++l) {
        const float *E = w + OFF_ENC + l * ENC_SIZE;
        net.off_in[l] = (int)pk.size();  pack_plain(pk, E + ENC_INW, 150, HID, TNH_150);
        net.off_out[l] = (int)pk.size(); pack_plain(pk, E + ENC_OW, HID, HID, TNH_50);
        net.off_f1[l] = (int)pk.size();  pack_plain(pk, E + ENC_L1W, 150, HID, TNH_150);
        net.off_f2[l] = (int)pk.size();  pack_plain(pk, E + ENC_L2W, HID, 150, TNH_50);
    }
    net.off_a0 = (int)pk.size(); pack_plain(pk, w + OFF_A0W, 100, 56, TNH_100);
    net.off_a2 = (int)pk.size(); pack_plain(pk, w + OFF_A2W, HID, 100, TNH_50);
    if (pk.size() > h->packed_floats) {
        cudaFree(h->packed_d);
        h->packed_d = nullptr;
        AEM_CUDA_CHECK(cudaMalloc(&h->packed_d, sizeof(float) * pk.size()));
        h->packed_floats = pk.size();
    }
    AEM_CUDA_CHECK(cudaDeviceSynchronize());   // no kernel may still read the old copy
    AEM_CUDA_CHECK(cudaMemcpy(h->raw_d, w, sizeof(float) * NUM_WEIGHTS, cudaMemcpyHostToDevice));
    AEM_CUDA_CHECK(cudaMemcpy(h->packed_d, pk.data(), sizeof(float) * pk.size(), cudaMemcpyHostToDevice));
    net.raw = h->raw_d;
    net.packed = h->packed_d;
    {   // tensor-core images + parameter block
        TcPack P;
        pack_tc(w, P);
        net.tc_pb_ln = (int)P.pblk.size();
        for (int l = 0; l < 3; ++l) {
            const float *E = w + OFF_ENC + l * ENC_SIZE;
            P.pblk.insert(P.pblk.end(), E + ENC_N1G, E + ENC_N1G + 200);    // n1g n1b n2g n2b are contiguous
        }
        net.tc_pb_pw = (int)P.pblk.size();
        P.pblk.insert(P.pblk.end(), w + OFF_PW, w + OFF_PW + 51);           // ponder weight (50) + bias
        net.tc_pb_a4 = (int)P.pblk.size();
        P.pblk.insert(P.pblk.end(), w + OFF_A4W, w + OFF_A4W + 51);         // last action layer (50) + bias
        while (P.pblk.size() % 4) P.pblk.push_back(0.0f);
        if (P.pblk.size() > (size_t)TC_PBLK_FLOATS) return aem_set_error(AEM_ERR_INVALID, "tc parameter block too large");
        if (P.img.size() > h->tc_img_floats) {
            cudaFree(h->tc_img_d);
            h->tc_img_d = nullptr;
            AEM_CUDA_CHECK(cudaMalloc(&h->tc_img_d, sizeof(float) * P.img.size()));
            h->tc_img_floats = P.img.size();
        }
        if (!h->tc_pblk_d) AEM_CUDA_CHECK(cudaMalloc(&h->tc_pblk_d, sizeof(float) * TC_PBLK_FLOATS));
        AEM_CUDA_CHECK(cudaMemcpy(h->tc_img_d, P.img.data(), sizeof(float) * P.img.size(), cudaMemcpyHostToDevice));
        AEM_CUDA_CHECK(cudaMemcpy(h->tc_pblk_d, P.pblk.data(), sizeof(float) * P.pblk.size(), cudaMemcpyHostToDevice));
        net.tc_img = h->tc_img_d;
        net.tc_pblk = h->tc_pblk_d;
        net.tc_pblk_floats = (int)P.pblk.size();
        if (!h->tc_layers_d) AEM_CUDA_CHECK(cudaMalloc(&h->tc_layers_d, sizeof(TcLayer) * TL_NUM));
        AEM_CUDA_CHECK(cudaMemcpy(h->tc_layers_d, P.layers, sizeof(TcLayer) * TL_NUM, cudaMemcpyHostToDevice));
        net.tc_layers_d = h->tc_layers_d;
    }
    {   // FP16x3 images
        H16Pack P;
        pack_h16(w, P);
        if (P.pblk.size() != (size_t)H16_PBLK_FLOATS) return aem_set_error(AEM_ERR_INVALID, "h16 parameter block size");
        if (P.img.size() > h->h16_img_halves) {
            cudaFree(h->h16_img_d);
            h->h16_img_d = nullptr;
            AEM_CUDA_CHECK(cudaMalloc(&h->h16_img_d, sizeof(__half) * P.img.size()));
            h->h16_img_halves = P.img.size();
        }
        if (!h->h16_pblk_d) AEM_CUDA_CHECK(cudaMalloc(&h->h16_pblk_d, sizeof(float) * H16_PBLK_FLOATS));
        if (!h->h16_layers_d) AEM_CUDA_CHECK(cudaMalloc(&h->h16_layers_d, sizeof(TcLayer) * TL_NUM));
        AEM_CUDA_CHECK(cudaMemcpy(h->h16_img_d, P.img.data(), sizeof(__half) * P.img.size(), cudaMemcpyHostToDevice));
        AEM_CUDA_CHECK(cudaMemcpy(h->h16_pblk_d, P.pblk.data(), sizeof(float) * H16_PBLK_FLOATS, cudaMemcpyHostToDevice));
        AEM_CUDA_CHECK(cudaMemcpy(h->h16_layers_d, P.layers, sizeof(TcLayer) * TL_NUM, cudaMemcpyHostToDevice));
        net.h16_img = h->h16_img_d;
        net.h16_pblk = h->h16_pblk_d;
        net.h16_pblk_floats = H16_PBLK_FLOATS;
        net.h16_layers_d = h->h16_layers_d;
    }
    h->weights_set = true;
    return AEM_OK;
}

int aem_set_action_space(aem_handle h, const double *actions, int A)
{
    if (!h || !actions) return aem_set_error(AEM_ERR_INVALID, "null argument");
    if (A < 1 || A > AEM_MAX_ACTIONS) return aem_set_error(AEM_ERR_INVALID, "A must be in [1, 128]");
    if (h->actions_set && h->A == A && memcmp(h->actions_h, actions, sizeof(double) * 2 * A) == 0)
        return AEM_OK;                     // the same table again (serial callers set it before every step): nothing to upload
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(cudaDeviceSynchronize());
    memcpy(h->actions_h, actions, sizeof(double) * 2 * A);
    AEM_CUDA_CHECK(cudaMemcpy(h->actions_d, actions, sizeof(double) * 2 * A, cudaMemcpyHostToDevice));
    h->A = A;
    h->actions_set = true;
    return AEM_OK;
}

int aem_set_env_params(aem_handle h, const aem_env_params *p)
{
    if (!h || !p) return aem_set_error(AEM_ERR_INVALID, "null argument");
    if (!(p->time_step > 0) || !(p->human_timestep >= 0))
        return aem_set_error(AEM_ERR_INVALID, "time_step must be positive");
    h->params = *p;
    h->params_set = true;
    return AEM_OK;
}

int aem_set_act_mode(aem_handle h, int act_fixed, int act_steps)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (act_fixed && act_steps < 1) return aem_set_error(AEM_ERR_INVALID, "act_steps must be >= 1");
    h->net.act_fixed = act_fixed ? 1 : 0;
    h->net.act_steps = act_steps;
    return AEM_OK;
}

int aem_set_precision(aem_handle h, int precision)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (precision < 0 || precision > 3)
        return aem_set_error(AEM_ERR_INVALID, "precision must be 0 (fp32), 1 (tf32x3), 2 (fp16x3) or 3 (fp16x1: single pass, opt-in)");
    h->precision = precision;
    return AEM_OK;
}

int aem_orca(aem_handle h, int B, int n, const double *humans_d, const double *robot_d, double *new_vel_d,
             void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!h->params_set) return aem_set_error(AEM_ERR_STATE, "aem_set_env_params has not been called");
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (n - 1 + (h->params.robot_visible ? 1 : 0) > 32) return aem_set_error(AEM_ERR_INVALID, "too many neighbours");
    if (B == 0) return AEM_OK;      // empty batch: nothing to do (buffers may be null)
    if (!humans_d || !robot_d || !new_vel_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_orca(B, n, humans_d, robot_d, h->params, new_vel_d, as_stream(stream), h->nact_d));
    h->launches += (B > 0);
    return AEM_OK;
}

int aem_env_lookahead(aem_handle h, int B, int n, const double *robot_d, const double *humans_d,
                      const double *new_vel_d, const double *global_time_d, double *rewards_d, int32_t *info_d,
                      double *dmin_d, double *humans_next_d, void *stream)
{
    int rc = check_ready(h, false, true);
    if (rc) return rc;
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_d || !new_vel_d || !global_time_d || !rewards_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_env_lookahead(B, n, h->A, robot_d, humans_d, new_vel_d, global_time_d, h->actions_d, 0,
                                        h->lut_d, h->params, action_amax(h), rewards_d, info_d, dmin_d,
                                        humans_next_d, as_stream(stream), h->nact_d));
    h->launches += (B > 0);
    return AEM_OK;
}

int aem_lookahead(aem_handle h, int B, int n, const double *robot_d, const double *humans_next_d,
                  const double *rewards_d, double gamma_bar, double *values_d, float *net_values_d, int32_t *best_d,
                  uint8_t *act_steps_d, void *stream)
{
    int rc = check_ready(h, true, true);
    if (rc) return rc;
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_next_d || !rewards_d || !values_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if ((long long)B * h->A > 0x7fffffffLL) return aem_set_error(AEM_ERR_INVALID, "B * A overflows int32");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    LookaheadArgs a;
    memset(&a, 0, sizeof(a));
    a.net = h->net;
    a.mode = 0;
    a.n = n; a.T = n + 1;
    a.S = 128 / a.T; if (a.S > 32) a.S = 32;
    a.A = h->A;
    a.total_samples = B * h->A;
    a.dt = h->params.time_step;
    a.gamma_bar = gamma_bar;
    a.robot = robot_d; a.humans_next = humans_next_d; a.actions = h->actions_d; a.rewards = rewards_d;
    a.values = values_d; a.net_values = net_values_d; a.steps = act_steps_d;
    a.n_active = h->nact_d;
    if (h->nact_d && h->precision == 1)
        return aem_set_error(AEM_ERR_STATE, "per-environment human counts need precision 0 (fp32) or 2 (fp16x3)");
    AEM_CUDA_CHECK(h->precision >= 2 ? launch_lookahead_h16(h, a, as_stream(stream), h->precision == 3)
                   : h->precision == 1 ? launch_lookahead_tc(h, a, as_stream(stream))
                                       : launch_lookahead(h, a, as_stream(stream)));
    if (best_d) AEM_CUDA_CHECK(launch_select(values_d, robot_d, B, h->A, best_d, as_stream(stream)));
    h->launches += (best_d ? 2 : 1) * (B > 0);
    return AEM_OK;
}

int aem_env_apply(aem_handle h, int B, int n, double *robot_d, double *humans_d, const double *new_vel_d,
                  double *global_time_d, const int32_t *chosen_d, const double *rewards_d, const int32_t *info_d,
                  const double *dmin_d, double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d,
                  double *out_dmin_d, void *stream)
{
    int rc = check_ready(h, false, true);
    if (rc) return rc;
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_d || !new_vel_d || !global_time_d || !chosen_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_env_apply(B, n, h->A, robot_d, humans_d, new_vel_d, global_time_d, h->actions_d, 0, chosen_d,
                                    rewards_d, info_d, dmin_d, h->params.time_step, out_reward_d, out_done_d,
                                    out_info_d, out_dmin_d, as_stream(stream), h->nact_d));
    h->launches += (B > 0);
    return AEM_OK;
}

int aem_env_prepare(aem_handle h, int B, int n, const double *robot_d, const double *humans_d, const double *global_time_d,
                    double *new_vel_d, double *rewards_d, int32_t *info_d, double *dmin_d, double *humans_next_d, void *stream)
{
    int rc = check_ready(h, false, true);
    if (rc) return rc;
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (n - 1 + (h->params.robot_visible ? 1 : 0) > 32) return aem_set_error(AEM_ERR_INVALID, "too many neighbours");
    if (!robot_d || !humans_d || !new_vel_d || !global_time_d || !rewards_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_env_lookahead(B, n, h->A, robot_d, humans_d, nullptr, global_time_d, h->actions_d, 0, h->lut_d,
                                        h->params, action_amax(h), rewards_d, info_d, dmin_d, humans_next_d,
                                        as_stream(stream), h->nact_d, new_vel_d));
    h->launches += 1;
    return AEM_OK;
}

int aem_env_finish(aem_handle h, int B, int n, const double *values_d, double *robot_d, double *humans_d,
                   const double *new_vel_d, double *global_time_d, const double *rewards_d, const int32_t *info_d,
                   const double *dmin_d, int32_t *best_d, double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d,
                   double *out_dmin_d, const double *pool_robot_d, const double *pool_humans_d, int P, int32_t *case_idx_d,
                   int case_stride, unsigned long long *outcome_counts_d, void *stream)
{
    int rc = check_ready(h, false, true);
    if (rc) return rc;
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!values_d || !robot_d || !humans_d || !new_vel_d || !global_time_d || !rewards_d || !info_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (pool_robot_d && (!pool_humans_d || !case_idx_d || P < 1)) return aem_set_error(AEM_ERR_INVALID, "bad scenario pool");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_env_finish(B, n, h->A, values_d, robot_d, humans_d, new_vel_d, global_time_d, h->actions_d, 0,
                                     rewards_d, info_d, dmin_d, h->params.time_step, best_d, out_reward_d, out_done_d,
                                     out_info_d, out_dmin_d, h->nact_d, pool_robot_d, pool_humans_d, P, case_idx_d,
                                     case_stride, outcome_counts_d, h->pool_nact_d, as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_env_step_xy(aem_handle h, int B, int n, double *robot_d, double *humans_d, const double *new_vel_d,
                    double *global_time_d, const double *actions_d, double max_speed, double *scratch_d,
                    double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d, double *out_dmin_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!h->params_set) return aem_set_error(AEM_ERR_STATE, "aem_set_env_params has not been called");
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_d || !new_vel_d || !global_time_d || !actions_d || !scratch_d || !out_reward_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    // scratch: rewards [B] f64 | dmin [B] f64 | info [B] i32 of the one action of every environment
    double *rew = scratch_d, *dmn = scratch_d + B;
    int32_t *inf = reinterpret_cast<int32_t *>(scratch_d + 2 * (size_t)B);
    AEM_CUDA_CHECK(launch_env_lookahead(B, n, 1, robot_d, humans_d, new_vel_d, global_time_d, actions_d, 2, h->lut_d,
                                        h->params, max_speed, rew, inf, dmn, nullptr, as_stream(stream), h->nact_d));
    AEM_CUDA_CHECK(launch_env_apply(B, n, 1, robot_d, humans_d, new_vel_d, global_time_d, actions_d, 2, nullptr, rew, inf,
                                    dmn, h->params.time_step, out_reward_d, out_done_d, out_info_d, out_dmin_d,
                                    as_stream(stream), h->nact_d));
    h->launches += 2;
    return AEM_OK;
}

int aem_env_reset_done(aem_handle h, int B, int n, double *robot_d, double *humans_d, double *global_time_d,
                       const int32_t *done_d, const int32_t *info_d, const double *pool_robot_d,
                       const double *pool_humans_d, int P, int32_t *case_idx_d, int case_stride,
                       unsigned long long *outcome_counts_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS || P < 1) return aem_set_error(AEM_ERR_INVALID, "bad B / n / P");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_d || !global_time_d || !done_d || !pool_robot_d || !pool_humans_d || !case_idx_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_env_reset(B, n, robot_d, humans_d, global_time_d, done_d, info_d, pool_robot_d,
                                    pool_humans_d, P, case_idx_d, case_stride, outcome_counts_d, as_stream(stream),
                                    h->nact_d, h->pool_nact_d));
    h->launches += (B > 0);
    return AEM_OK;
}

int aem_value_forward(aem_handle h, int S, int n, const float *states_d, float *values_d, uint8_t *steps_d,
                      void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!h->weights_set) return aem_set_error(AEM_ERR_STATE, "aem_load_weights has not been called");
    if (S < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad S / n");
    if (S == 0) return AEM_OK;
    if (!states_d || !values_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    LookaheadArgs a;
    memset(&a, 0, sizeof(a));
    a.net = h->net;
    a.mode = 1;
    a.n = n; a.T = n + 1;
    a.S = 128 / a.T; if (a.S > 32) a.S = 32;
    a.A = 1;
    a.total_samples = S;
    a.states = states_d;
    a.net_values = values_d; a.steps = steps_d;
    AEM_CUDA_CHECK(h->precision >= 2 ? launch_lookahead_h16(h, a, as_stream(stream), h->precision == 3)
                   : h->precision == 1 ? launch_lookahead_tc(h, a, as_stream(stream))
                                       : launch_lookahead(h, a, as_stream(stream)));
    h->launches += (S > 0);
    return AEM_OK;
}

int aem_transform(aem_handle h, int B, int n, const double *robot_d, const double *humans_d, float *states_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (B < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (B == 0) return AEM_OK;
    if (!robot_d || !humans_d || !states_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_transform(robot_d, humans_d, B, n, states_d, as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_gru_act(aem_handle h, int S, int T, const float *x_d, float *out_d, uint8_t *steps_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!h->weights_set) return aem_set_error(AEM_ERR_STATE, "aem_load_weights has not been called");
    if (S < 0 || T < 1 || T > AEM_MAX_HUMANS + 1) return aem_set_error(AEM_ERR_INVALID, "bad S / T");
    if (S == 0) return AEM_OK;
    if (!x_d || !out_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    LookaheadArgs a;
    memset(&a, 0, sizeof(a));
    a.net = h->net;
    a.mode = 2;
    a.n = T - 1; a.T = T;
    a.S = 128 / T; if (a.S > 32) a.S = 32;
    a.A = 1;
    a.total_samples = S;
    a.states = x_d;
    a.gru_out = out_d; a.steps = steps_d;
    AEM_CUDA_CHECK(h->precision >= 2 ? launch_lookahead_h16(h, a, as_stream(stream), h->precision == 3)
                   : h->precision == 1 ? launch_lookahead_tc(h, a, as_stream(stream))
                                       : launch_lookahead(h, a, as_stream(stream)));
    h->launches += (S > 0);
    return AEM_OK;
}

static int ensure_scratch(aem_handle h, int B, int n)
{
    if (B <= h->cap_B && n <= h->cap_n) return AEM_OK;
    const int cB = B > h->cap_B ? B : h->cap_B, cn = n > h->cap_n ? n : h->cap_n;
    cudaFree(h->s_robot); cudaFree(h->s_humans); cudaFree(h->s_time); cudaFree(h->s_newvel);
    cudaFree(h->s_rewards); cudaFree(h->s_dmin); cudaFree(h->s_hnext); cudaFree(h->s_values);
    cudaFree(h->s_oreward); cudaFree(h->s_info); cudaFree(h->s_best); cudaFree(h->s_odone); cudaFree(h->s_oinfo);
    h->cap_B = h->cap_n = 0;
    const size_t BA = (size_t)cB * AEM_MAX_ACTIONS;
    AEM_CUDA_CHECK(cudaMalloc(&h->s_robot, sizeof(double) * cB * 9));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_humans, sizeof(double) * cB * cn * 8));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_time, sizeof(double) * cB));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_newvel, sizeof(double) * cB * cn * 2));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_rewards, sizeof(double) * BA));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_dmin, sizeof(double) * BA));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_hnext, sizeof(double) * cB * cn * 5));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_values, sizeof(double) * BA));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_oreward, sizeof(double) * cB));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_info, sizeof(int32_t) * BA));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_best, sizeof(int32_t) * cB));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_odone, sizeof(int32_t) * cB));
    AEM_CUDA_CHECK(cudaMalloc(&h->s_oinfo, sizeof(int32_t) * cB));
    if (!h->s_stream) {
        cudaStream_t st;
        AEM_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        h->s_stream = st;
    }
    h->cap_B = cB; h->cap_n = cn;
    return AEM_OK;
}

int aem_decide_step_host(aem_handle h, int B, int n, double gamma_bar, double *robot_h, double *humans_h,
                         double *global_time_h, int32_t *chosen_h, double *reward_h, int32_t *done_h,
                         int32_t *info_h, double *values_h)
{
    int rc = check_ready(h, true, true);
    if (rc) return rc;
    if (h->nact_d) return aem_set_error(AEM_ERR_STATE, "aem_decide_step_host takes uniform batches: clear aem_set_active_counts first");
    if (B < 1 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (!robot_h || !humans_h || !global_time_h || !chosen_h || !reward_h || !done_h || !info_h)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    rc = ensure_scratch(h, B, n);
    if (rc) return rc;
    cudaStream_t st = as_stream(h->s_stream);
    const int A = h->A;
    AEM_CUDA_CHECK(cudaMemcpyAsync(h->s_robot, robot_h, sizeof(double) * B * 9, cudaMemcpyHostToDevice, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(h->s_humans, humans_h, sizeof(double) * B * n * 8, cudaMemcpyHostToDevice, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(h->s_time, global_time_h, sizeof(double) * B, cudaMemcpyHostToDevice, st));
    // three launches: ORCA + env lookahead, the value lookahead, argmax + env step
    rc = aem_env_prepare(h, B, n, h->s_robot, h->s_humans, h->s_time, h->s_newvel, h->s_rewards, h->s_info, h->s_dmin,
                         h->s_hnext, st);
    if (rc) return rc;
    rc = aem_lookahead(h, B, n, h->s_robot, h->s_hnext, h->s_rewards, gamma_bar, h->s_values, nullptr, nullptr, nullptr, st);
    if (rc) return rc;
    rc = aem_env_finish(h, B, n, h->s_values, h->s_robot, h->s_humans, h->s_newvel, h->s_time, h->s_rewards, h->s_info,
                        h->s_dmin, h->s_best, h->s_oreward, h->s_odone, h->s_oinfo, nullptr, nullptr, nullptr, 0, nullptr, 0,
                        nullptr, st);
    if (rc) return rc;
    AEM_CUDA_CHECK(cudaMemcpyAsync(robot_h, h->s_robot, sizeof(double) * B * 9, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(humans_h, h->s_humans, sizeof(double) * B * n * 8, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(global_time_h, h->s_time, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(chosen_h, h->s_best, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(reward_h, h->s_oreward, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(done_h, h->s_odone, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaMemcpyAsync(info_h, h->s_oinfo, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    if (values_h)
        AEM_CUDA_CHECK(cudaMemcpyAsync(values_h, h->s_values, sizeof(double) * B * A, cudaMemcpyDeviceToHost, st));
    AEM_CUDA_CHECK(cudaStreamSynchronize(st));
    return AEM_OK;
}

int aem_generate_scenarios(aem_handle h, int rule, uint32_t first_seed, int count, int n, const aem_scenario_params *p,
                           double *robot_d, double *humans_d, void *stream)
{
    if (!h || !p) return aem_set_error(AEM_ERR_INVALID, "null argument");
    if (rule != 0 && rule != 1) return aem_set_error(AEM_ERR_INVALID, "rule must be 0 (circle_crossing) or 1 (square_crossing)");
    if (count < 0 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad count / n");
    if (count == 0) return AEM_OK;
    if (!robot_d || !humans_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_scenarios(rule, first_seed, count, n, *p, robot_d, humans_d, as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_generate_mixed_scenarios(aem_handle h, uint32_t first_seed, int count, const aem_scenario_params *p, double *robot_d,
                                 double *humans_d, int32_t *n_active_d, void *stream)
{
    if (!h || !p) return aem_set_error(AEM_ERR_INVALID, "null argument");
    if (count < 0) return aem_set_error(AEM_ERR_INVALID, "bad count");
    if (count == 0) return AEM_OK;
    if (!robot_d || !humans_d || !n_active_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_scenarios(2, first_seed, count, 5, *p, robot_d, humans_d, as_stream(stream), n_active_d));
    h->launches += 1;
    return AEM_OK;
}

int aem_train_grad(aem_handle h, int B, int n, const float *params_d, const float *states_d, const float *targets_d,
                   float *grad_d, float *loss_d, float *values_d, int32_t *steps_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (B < 1 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad B / n");
    if (!params_d || !states_d || !targets_d || !grad_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (((uintptr_t)params_d & 15) != 0) return aem_set_error(AEM_ERR_INVALID, "params_d must be 16-byte aligned (bulk copies)");
    if (h->net.act_fixed && h->net.act_steps > 3) return aem_set_error(AEM_ERR_INVALID, "aem_train_grad: act_fixed with at most 3 steps");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_train_grad(h, B, n, params_d, states_d, targets_d, grad_d, loss_d, values_d, steps_d, as_stream(stream)));
    h->launches += 3;
    return AEM_OK;
}

int aem_set_active_counts(aem_handle h, int32_t *n_active_d, const int32_t *pool_n_active_d)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    h->nact_d = n_active_d;
    h->pool_nact_d = pool_n_active_d;
    return AEM_OK;
}

int aem_set_train_dropout(aem_handle h, double p, uint64_t seed)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!(p >= 0.0 && p < 1.0)) return aem_set_error(AEM_ERR_INVALID, "dropout probability must be in [0, 1)");
    h->drop_p = p;
    h->drop_seed = seed;
    h->drop_ctr = 0;
    return AEM_OK;
}

int aem_sample_batch(aem_handle h, int N, int k, int n, const float *states_d, const float *values_d, uint64_t seed,
                     uint64_t counter, float *out_states_d, float *out_values_d, int64_t *out_index_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (N < 1 || k < 1 || k > N || k > 8192 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad N / k / n");
    if (!states_d || !values_d || !out_states_d || !out_values_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_sample_batch(N, k, n * 13, states_d, values_d, seed, counter, out_states_d, out_values_d, out_index_d,
                                       as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_adam_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *exp_avg_d, float *exp_avg_sq_d,
                  double lr, double beta1, double beta2, double eps, int step, double grad_scale, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!params_d || !grad_d || !exp_avg_d || !exp_avg_sq_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (step < 1 || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1)) return aem_set_error(AEM_ERR_INVALID, "bad step / betas");
    if (count == 0) return AEM_OK;
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_adam(count, params_d, grad_d, exp_avg_d, exp_avg_sq_d, lr, beta1, beta2, eps, step, grad_scale,
                               as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_train_step_adam(aem_handle h, int N, int k, int n, const float *ring_states_d, const float *ring_values_d, uint64_t seed,
                        unsigned long long *dev_state_d, float *adam_scal_d, float *params_d, float *grad_d, float *exp_avg_d,
                        float *exp_avg_sq_d, double lr, double beta1, double beta2, double eps, float *batch_states_d,
                        float *batch_values_d, float *loss_d, float *loss_sum_d, int32_t *steps_d, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (N < 1 || k < 1 || k > N || k > 8192 || n < 1 || n > AEM_MAX_HUMANS) return aem_set_error(AEM_ERR_INVALID, "bad N / k / n");
    if (!ring_states_d || !ring_values_d || !dev_state_d || !adam_scal_d || !params_d || !grad_d || !exp_avg_d || !exp_avg_sq_d ||
        !batch_states_d || !batch_values_d)
        return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (((uintptr_t)params_d & 15) != 0) return aem_set_error(AEM_ERR_INVALID, "params_d must be 16-byte aligned (bulk copies)");
    if (h->net.act_fixed && h->net.act_steps > 3) return aem_set_error(AEM_ERR_INVALID, "aem_train_step_adam: act_fixed with at most 3 steps");
    if (!(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1)) return aem_set_error(AEM_ERR_INVALID, "bad betas");
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_train_step_adam(h, N, k, n, ring_states_d, ring_values_d, seed, dev_state_d, adam_scal_d, params_d, grad_d,
                                          exp_avg_d, exp_avg_sq_d, lr, beta1, beta2, eps, batch_states_d, batch_values_d, loss_d,
                                          loss_sum_d, steps_d, as_stream(stream)));
    h->launches += 5;
    return AEM_OK;
}

int aem_sgd_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *momentum_buf_d, double lr,
                 double momentum, int step, double grad_scale, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!params_d || !grad_d || !momentum_buf_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (step < 1) return aem_set_error(AEM_ERR_INVALID, "step is 1-based");
    if (count == 0) return AEM_OK;
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_sgd(count, params_d, grad_d, momentum_buf_d, lr, momentum, step, grad_scale, as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

int aem_rmsprop_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *square_avg_d, double lr,
                     double alpha, double eps, double grad_scale, void *stream)
{
    if (!h) return aem_set_error(AEM_ERR_INVALID, "null handle");
    if (!params_d || !grad_d || !square_avg_d) return aem_set_error(AEM_ERR_INVALID, "null buffer");
    if (count == 0) return AEM_OK;
    AEM_CUDA_CHECK(cudaSetDevice(h->device));
    AEM_CUDA_CHECK(launch_rmsprop(count, params_d, grad_d, square_avg_d, lr, alpha, eps, grad_scale, as_stream(stream)));
    h->launches += 1;
    return AEM_OK;
}

long long aem_launch_count(aem_handle h) { return h ? h->launches : 0; }
long long aem_train_scratch_generation(aem_handle h) { return h ? h->tr_generation : 0; }

}  // extern "C"
