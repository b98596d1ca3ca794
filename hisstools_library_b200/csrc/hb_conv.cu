// hb_conv.cu -- host side of the uniform partitioned convolution engine behind the hb_conv_* entry
// points of include/hisstools_b200.h (kernels: hb_conv_kernels.cuh).
//
// Streaming bookkeeping (what PartitionedConvolve.cpp:243-385 does with its four FFT buffers and the
// RW counter) is kept on linear device staging rows instead of rings:
//   xin row  = [previous hop: B samples][pending samples of the current hop: rw][new samples of this call]
//   yout row = [result block of the last completed hop: B][blocks of the hops completed in this call]
// so hop h of a call transforms xin[h*B .. h*B + 2B) and the n output samples of the call are
// yout[rw .. rw + n): the output is the linear convolution delayed by exactly B for any call sizes
// (SURVEY A.2).  After a call the tail of both rows becomes the head of the other (ping-pong) row set.
#include "hb_common.cuh"
#include "hb_conv_kernels.cuh"
#include "hb_conv_big.cuh"
#include "hb_conv_cluster.cuh"
#include "hb_conv_fused.cuh"
#include "hb_conv_mh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

using namespace hb;

namespace
{
constexpr uintptr_t MIN_FFT_LOG2 = 5;       // PartitionedConvolve.h:18
constexpr uintptr_t MAX_FFT_LOG2 = 20;      // PartitionedConvolve.h:19
// inbox tail: arrival counters [HB_INBOX_DEPTH][world] (<= 1024 bytes), then the late-peer flag word
constexpr size_t INBOX_LATE_OFF = 1024, INBOX_TAIL_BYTES = 1280;
constexpr uint32_t MH_MAX = 8;              // hops one multi-hop multiply-accumulate launch covers at most
constexpr uint32_t MH_EXTRA = MH_MAX - 1;   // extra delay-line slots that needs

// reference error codes (ConvolveErrors.h:4-19)
enum
{
    ERR_NONE = 0,
    ERR_IN_CHAN = 1,
    ERR_OUT_CHAN = 2,
    ERR_MEM_UNAVAILABLE = 3,
    ERR_MEM_ALLOC_TOO_SMALL = 4,
    ERR_TIME_IMPULSE_TOO_LONG = 5,
    ERR_TIME_LENGTH_OUT_OF_RANGE = 6,
    ERR_PARTITION_LENGTH_TOO_LARGE = 7,
    ERR_FFT_SIZE_MAX_TOO_SMALL = 8,
    ERR_FFT_SIZE_MAX_TOO_LARGE = 9,
    ERR_FFT_SIZE_MAX_NON_POWER_OF_TWO = 10,
    ERR_FFT_SIZE_OUT_OF_RANGE = 11,
    ERR_FFT_SIZE_NON_POWER_OF_TWO = 12
};

// ceil(log2(value)) as PartitionedConvolve::log2 (PartitionedConvolve.cpp:114-129)
uintptr_t ceil_log2(uintptr_t value)
{
    uintptr_t bits = 0;
    for (uintptr_t v = value; v; v >>= 1) bits++;
    if (!bits) return 0;
    return value == (uintptr_t(1) << (bits - 1)) ? bits - 1 : bits;
}
} // namespace

struct hb_conv
{
    int dtype = HB_F32;
    int device = 0;
    uint32_t groups = 1, ins = 1, outs = 1;
    uintptr_t max_fft_log2 = 0, fft_log2 = 0;
    uintptr_t max_length = 0;       // taps per pair, rounded up to a multiple of max_fft/2 (cpp:77-82)
    uintptr_t offset = 0, length = 0;
    intptr_t reset_offset = -1;
    bool need_reset = true;

    // device memory
    void *d_H = nullptr;            // IR spectra
    void *d_X = nullptr;            // FDL
    void *d_Hnyq = nullptr, *d_Xnyq = nullptr;
    DevBuf d_S;                     // stream-K partial segments
    void *d_tw = nullptr;
    int tw_log2 = 1;
    DevBuf d_xin[2], d_yout[2];     // ping-pong staging rows
    size_t xin_ld = 0, yout_ld = 0;
    int cur = 0;                    // which staging set holds the retained state
    size_t x_tail = 0, y_tail = 0;  // where the retained head starts inside set `cur`
    DevBuf d_io_in, d_io_out, d_ir; // host-call staging on the device
    PinnedBuf h_in, h_out, h_ir;
    cudaStream_t stream = nullptr;

    // streaming state
    uintptr_t rw = 0;               // samples already received of the current hop
    std::vector<uint32_t> nparts;   // partitions loaded per pair [group][out][in]
    uint32_t P = 0;                 // max over pairs = ring length
    Geom g{};
    int sm_count = 148;
    int ctas_per_sm = 1;
    int variant = 1;                // 1 = TMA ring, 0 = direct loads
    int nstages = 0;
    size_t cmac_smem = 0;
    std::mutex lock;

    // schedule of a hop (hb_conv_set_schedule).  Overlapped: partitions 1..P-1 only meet spectra that are
    // already in the delay line, so their share of hop t+1 ("tail") is computed on a second stream as soon as
    // the forward FFTs of hop t are done, beside the inverse FFTs of hop t and the forward FFTs of hop t+1; the
    // critical path of a hop is then forward FFT -> partition 0 ("head") -> inverse FFT.  This is the load
    // spreading of PartitionedConvolve.cpp:330-347 (partitions done ahead, between hops) in stream form.
    int schedule = 2;               // requested: 0 serial, 1 overlapped, 2 automatic, 3 fused (one cluster launch per hop) where eligible
    bool split = false;             // overlapped schedule in effect for the current geometry (needs P >= 2)
    bool fused = false;             // fused single-launch hop in effect (hb_conv_fused.cuh)
    uint32_t fused_cs = 1;          // its cluster size
    // consecutive fused hops that overlap (hb_conv_fused.cuh, "chained"; hb_conv_set_hop_overlap)
    int hop_overlap = 1;            // 0 never, 1 on the engine's own stream, 2 on any stream (the caller's rows are complete when a call is made)
    DevBuf d_chain;                 // the three counters of the hop chain (+ padding)
    unsigned long long chain_n = 0; // fused hops launched since the counters were zeroed
    bool chain_ok = false;          // the last thing this engine enqueued was a fused hop, on chain_stream
    bool last_hop_chained = false;  // the last fused hop launched ran in the chained order (it did not wait for its predecessor up front)
    cudaStream_t chain_stream = nullptr;
    struct ByteRange { const char *b = nullptr, *e = nullptr; };
    ByteRange chain_out[16];        // rows the last fused hops write (block handed over + block kept): a hop whose input rows touch one of
    uint32_t chain_out_pos = 0;     // them is fed by a hop that may still be running and keeps the strict order
    Range r_full{}, r_head{}, r_tail{};
    DevBuf d_St[2];                 // tail partial segments, double-buffered over hops
    cudaStream_t s_tail = nullptr, s_tail_b = nullptr;
    int tail_streams_req = 0;       // hb_conv_set_tail_streams: 0 automatic, 1, or 2 = the tails of consecutive hops alternate between two
    int tail_streams = 1;           // streams, so that the CTAs of the next tail take over the SMs as the CTAs of the running one exit (in effect)
    cudaEvent_t ev_fwd = nullptr, ev_tail[2] = {nullptr, nullptr};
    bool tail_valid = false;        // d_St[tail_par] holds the tail of the upcoming hop
    bool tail_missing = false;      // a multi-hop batch ran last: nothing was computed ahead for the upcoming hop
    int tail_par = 0;
    DevBuf d_trace;                 // optional kernel timeline (hb_conv_set_trace)
    DevBuf d_Smh;                   // partial segments of a multi-hop launch: [hop][cta + tile][row][TBV]
    uint32_t extra_slots = MH_EXTRA;  // delay-line slots beyond P: the hops one launch may look ahead (63 while the delay line stays small)
    uint32_t hb_max = 1;            // hops one batch of a launch-latency-bound engine covers at most (process_core), 1 = no batches
    int multi_hop = 1;              // hb_conv_set_multi_hop: batch the hops of one call over a single pass of the IR spectra
    bool mh_ok = false;             // eligible for the current geometry (plan_geometry)
    int mh_max = 1;                 // hops one multi-hop launch covers at most for the current geometry
    bool mh_packed = false;         // float engines: the packed-FFMA2 kernel (hb_conv_mh.cuh), up to 8 hops per pass
    int mh_stages = 3;
    int fft_path = 0;               // hb_conv_set_fft_path: 0 automatic, 1 one CTA per transform, 2 cluster of 8 CTAs, 3 four-step
    BigScratch big;                 // four-step scratch for FFT sizes above the single-CTA limit (hb_conv_big.cuh)
    DevBuf d_nyq;
    DevBuf d_part;                  // fused multi-GPU exchange at four-step sizes: this rank's partial blocks before delivery

    // deferred host-pointer path of hb_conv_process: the block finished by a hop is fetched to pinned host
    // memory while the caller is away; a later call only waits on an event that completed long ago
    bool deferred = true;
    PinnedBuf h_blk[2], h_inq[2];
    DevBuf d_inq[2];
    DevBuf d_blk[2];                // large finished blocks are packed on the device first: one contiguous download instead of a pitched one
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;      // copies run beside the hop kernels, not between them
    cudaEvent_t ev_blk[2] = {nullptr, nullptr}, ev_inq[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool done_pending[2] = {false, false};
    bool blk_pending[2] = {false, false}, inq_pending[2] = {false, false};
    int blk_cur = 0, inq_cur = 0;
    bool blk_valid = false;         // h_blk[blk_cur] holds the last completed block (outputs are read from it at offset rw)
    size_t blk_B = 0;

    // fused multi-GPU exchange (hb_conv_shard_*): this rank's inbox and the peers' inboxes opened over CUDA IPC
    uint32_t shard_world = 0, shard_rank = 0;
    void *d_inbox = nullptr;
    size_t inbox_data_bytes = 0, inbox_slot = 0;
    void *peer_base[HB_MAX_WORLD] = {};
    bool peers_attached = false;
    bool peers_local = false;       // peers of the same process (plain pointers, nothing to close)
    uint32_t hop_seq = 0, parity_uses[HB_INBOX_DEPTH] = {};

    // optional per-kernel timing (hb_conv_set_profiling): PROF_EV events per hop, five on the launching stream
    // (around the forward FFTs, the whole / head multiply-accumulate, the wait for the tail, the inverse FFTs)
    // and two on the tail stream around the tail multiply-accumulate
    bool profiling = false;
    std::vector<cudaEvent_t> ev;
    std::vector<char> ev_has_tail;
    size_t ev_used = 0;             // hops recorded and not yet drained
    double prof_ms[5] = {0, 0, 0, 0, 0};  // forward, whole/head multiply-accumulate, wait for tail, inverse, tail
    uint64_t prof_hops = 0;

    // pairs restarted while the stream runs (hb_conv_set_ir_live / hb_conv_reset_pair): their spectra wait in a private buffer and
    // are copied into the IR array one partition per hop (k_pair_copy)
    struct Reveal
    {
        uint32_t grp, in, out;
        uint32_t np;                // partitions of the pair
        uint32_t shown;             // partitions already in place
        uint32_t age;               // hops processed since the restart
        void *side;                 // [np][B] bins in natural order, then np Nyquist values
    };
    std::vector<Reveal> reveals;

    size_t esize() const { return dtype_size(dtype); }
    size_t pairs() const { return size_t(groups) * ins * outs; }
};

namespace
{

// largest half-length transform (log2 of complex points) run as one CTA in shared memory; above it the four-step
// chains of hb_conv_big.cuh take over.  HB_BIG_FROM (experiments only) lowers the switch-over, down to 2^12 points.
int single_cta_max_log2m(const hb_conv *c)
{
    const int lim = c->dtype == HB_F64 ? SmemFftLimit<double>::max_log2m : SmemFftLimit<float>::max_log2m;
    static const char *env = getenv("HB_BIG_FROM");
    if (env && atoi(env) >= 12) return std::min(lim, atoi(env) - 1);
    if (c->fft_path == 3) return std::min(lim, 11);
    return lim;
}

// transforms spread over clusters of 8 CTAs (hb_conv_cluster.cuh).  Automatic choice: double precision with fewer
// transforms than two per SM -- there the one-CTA kernels are latency-bound on the few SMs they occupy and set the hop period
// (config 5: 38 us forward, 60-86 us inverse on 16 SMs; profiles/r1_c5_cluster_fft.txt).
bool use_cluster_fft(const hb_conv *c)
{
    const int m = (int) c->g.log2n - 1;
    if (c->fft_path == 1 || c->fft_path == 3 || c->fused) return false;
    if (m < CL_MIN_LOG2M || m > single_cta_max_log2m(c) || c->g.n_bt > (uint32_t) CL_MAX_SEG_TILES) return false;
    if (c->fft_path == 2) return true;
    return c->dtype == HB_F64 && uint64_t(c->groups) * std::max(c->ins, c->outs) <= 2 * (uint64_t) c->sm_count;
}

// ---- geometry -------------------------------------------------------------------------------------
uint32_t choose_ot(uint32_t outs)
{
    uint32_t best = 1;
    uint64_t best_cost = ~uint64_t(0);
    for (uint32_t ot = 1; ot <= 64; ot <<= 1)
    {
        uint64_t cost = uint64_t((outs + ot - 1) / ot) * (ot + 1);
        if (cost <= best_cost) { best_cost = cost; best = ot; }
    }
    return best;
}

void plan_geometry(hb_conv *c)
{
    Geom &g = c->g;
    const uint32_t cpv = c->dtype == HB_F64 ? 1 : 2;
    g.groups = c->groups; g.ins = c->ins; g.outs = c->outs;
    g.log2n = (uint32_t) c->fft_log2;
    g.B = 1u << (g.log2n - 1);
    g.Pcap = (uint32_t) (c->max_length / g.B);
    g.P = c->P;
    g.R = c->P ? c->P + c->extra_slots : 0;       // the delay line keeps the look-ahead slots of a multi-hop launch
    g.OT = choose_ot(c->outs);
    g.n_ot = (c->outs + g.OT - 1) / g.OT;
    const uint32_t xq = g.B / cpv;
    g.TBV = std::min<uint32_t>(xq, std::max<uint32_t>(2048u / g.OT, 1u));
    g.n_bt = xq / g.TBV;
    g.TX = std::min<uint32_t>(g.TBV, 256u);
    g.XA = g.TBV / g.TX;
    g.TY = std::min<uint32_t>(g.OT, 256u / g.TX);
    g.OB = g.OT / g.TY;
    g.Q = g.OT * g.TBV;
    g.upt = g.ins * g.P;
    g.tiles = g.groups * g.n_ot * g.n_bt;
    g.U = uint64_t(g.tiles) * g.upt;
    const uint32_t log2m = g.log2n - 1;
    const uint64_t sms = (uint64_t) c->sm_count;
    // every tile is summed by ONE inverse-FFT CTA: with few tiles (few outputs, short hops) a wide grid would
    // hand that CTA hundreds of partial segments, so a grid is narrowed to at most 16 segments per tile
    auto make_range = [&](uint32_t p0, uint32_t pc, uint64_t want)
    {
        Range r{};
        r.p0 = p0; r.pc = pc;
        r.upt = g.ins * pc;
        r.U = uint64_t(g.tiles) * r.upt;
        want = std::min<uint64_t>(want, uint64_t(g.tiles) * 16);
        r.G = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(r.U, want));
        return r;
    };
    const uint64_t per_sm = c->variant == 1 ? 1 : std::max(1, c->ctas_per_sm);
    // automatic choice: below a few MiB of tail spectra a hop is bound by launch latency, and the extra launch and
    // the two stream hand-overs of the overlapped schedule cost more than they hide (measured: DESIGN.md 4)
    const uint64_t tail_bytes = uint64_t(c->pairs() + uint64_t(g.groups) * g.ins) * (g.P ? g.P - 1 : 0) * g.B * 2 * c->esize();
    // fused single-launch hop (hb_conv_fused.cuh): engines of up to 8 outputs whose partition spectrum is one bin tile and
    // whose per-output share of the delay line is small enough for one cluster (<= 8 CTAs x 256 KiB of L2-resident reads)
    c->fused = false;
    c->fused_cs = 1;
    if ((c->schedule == 2 || c->schedule == 3) && g.outs <= 8 && g.n_bt == 1 && g.P >= 1 && log2m <= 12 && log2m >= 3)
    {
        // spectra one output's cluster reads per hop: its IR partitions and the delay line of every input
        const uint64_t per_output = uint64_t(g.ins) * (g.P - 1) * g.B * 4 * c->esize();
        uint32_t cs = 1;
        static const char *env_cs = getenv("HB_FUSED_MAX_CS"), *env_kb = getenv("HB_FUSED_MAX_KB");      // experiments only
        const uint32_t cs_max = env_cs && atoi(env_cs) >= 1 ? (uint32_t) atoi(env_cs) : 16u;
        const uint64_t per_rank_max = (env_kb && atoi(env_kb) > 0 ? (uint64_t) atoi(env_kb) : 1152u) << 10;
        while (cs < cs_max && per_output / cs > (uint64_t(96) << 10)) cs <<= 1;
        // measured: config 1 15 us per hop against 21 for the three-kernel hop, config 2 21 against 27 (profiles/r1_small_hops.txt).  With
        // programmatic dependent launch (hb_conv_fused.cuh) all partitions >= 2 are multiplied beside the previous hop, so a rank may
        // read much more: config 3 (8 -> 1, 16 MiB of spectra per hop) on a cluster of 16 with 1 MiB per rank runs 21.8 us per hop against
        // 34.8 overlapped (and 41 us on a cluster of 8 with 2 MiB per rank); profiles/r2_small_hops.txt
        if (per_output / cs <= per_rank_max) { c->fused = true; c->fused_cs = cs; }
    }
    c->split = !c->fused && g.P >= 2 && (c->schedule == 1 || ((c->schedule == 2 || c->schedule == 3) && tail_bytes >= (uint64_t(4) << 20)));
    // Short tail launches lose a visible share of the hop to their ramp-up and drain (about 10 us per launch): with two tail streams
    // the next launch takes over the SMs as the running one leaves them.  Measured (profiles/r2_tail_streams.txt): one rank of config 4
    // on 8 GPUs (1.07 GB per launch) 167.9 -> 164.5 us per hop, config 5 (0.53 GB) 90.9 -> 87.0 us, config 4 on one GPU (8.6 GB) unchanged.
    c->tail_streams = c->tail_streams_req ? c->tail_streams_req : (tail_bytes <= (uint64_t(3) << 30) ? 2 : 1);
    // TMA ring depth: about 96 KB in flight per SM, 3 to 6 stages.  Measured on B200 at config 4 (32.5 KB stages,
    // profiles/r1_ring_depth.txt): 3 stages stream 7.2 TB/s, 2 stages 6.6, 5-6 stages 6.5 -- twice Little's law for
    // the chip (7.3 TB/s x ~1 us / 148 SMs = 49 KB) is enough, and a deeper ring only lowers the DRAM efficiency.
    const size_t stage = size_t(g.Q + g.TBV) * 16;
    int st = (int) std::max<size_t>(3, std::min<size_t>(6, (96 * 1024 + stage / 2) / stage));
    uint64_t reserve = 0;
    if (c->split)
    {
        // The tail launch runs beside the FFT kernels of the critical path.  Where an FFT CTA fits on an SM next
        // to a multiply-accumulate CTA and its ring (shared memory; registers: 512 x 64 + 256 x 96 for transforms
        // up to 4096 points, see HB_CMAC_MAXREG) nothing changes; otherwise the tail grid leaves as many SMs free
        // as the FFT kernels have CTAs (stream-K: any grid size balances).
        const size_t es = c->esize();
        const size_t fft_data = size_t(padded_elems<HB_PADSH>(1u << log2m)) * 2 * es;
        const size_t fft_tw = (size_t(1) << log2m) * 2 * es;
        const size_t fft_total = fft_data + (fft_data + fft_tw <= 200 * 1024 ? fft_tw : 0) + 3 * 1024;   // static + driver reserve
        const size_t budget = 227 * 1024 > fft_total + 2048 ? 227 * 1024 - fft_total - 2048 : 0;
        const int st_co = (int) (budget / stage);
        const uint64_t fft_ctas = uint64_t(g.groups) * std::max(g.ins, g.outs);
        // (the four-step and the cluster kernels are small CTAs: they fit beside the ring)
        if (c->variant == 1 && !(log2m <= 12 && st_co >= st) && (int) log2m <= single_cta_max_log2m(c) && !use_cluster_fft(c)) reserve = std::min<uint64_t>(fft_ctas, sms / 4);
    }
    // cluster FFT CTAs (up to 42 KiB) are placed beside the resident tail CTA: leave them room
    if (c->split && c->variant == 1 && use_cluster_fft(c)) st = std::min<int>(st, (int) ((227 * 1024 - 48 * 1024) / stage));
    static const char *env_st = getenv("HB_STAGES"), *env_rs = getenv("HB_RESERVE");          // experiments only
    if (env_st && atoi(env_st) >= 2) st = std::min<int>(atoi(env_st), (int) ((220 * 1024) / stage));
    if (env_rs && c->split) reserve = std::min<uint64_t>((uint64_t) atoi(env_rs), sms - 1);
    c->nstages = std::max(2, st);
    c->cmac_smem = size_t(c->nstages) * stage + size_t(c->nstages) * 8;
    // multi-hop reuse (k_cmac_tma_mh): HBM-bound engines with full-height tiles (the FDL tile is 1/OT of the IR unit, so the
    // extra hops add little traffic), TMA variant, single-CTA transforms
    {
        const size_t stage_mh = size_t(g.Q + 4 * g.TBV) * 16;              // scalar kernel (double): at most 4 hops per pass
        c->mh_stages = (int) std::min<size_t>(3, (200 * 1024) / stage_mh);
        c->mh_ok = c->multi_hop && c->variant == 1 && !c->fused && tail_bytes >= (uint64_t(4) << 20) && g.OT >= 8 &&
                   ((g.XA == 1 && g.OB == 8) || (g.XA == 2 && g.OB == 4)) && (int) log2m <= single_cta_max_log2m(c) &&
                   c->mh_stages >= 2;
        static const char *env_scalar = getenv("HB_MH_SCALAR");              // experiments only: the scalar kernel on float engines
        c->mh_packed = c->mh_ok && c->dtype == HB_F32 && mh2_supported(g, 2) && !(env_scalar && atoi(env_scalar));
        // batches of hops on engines that are NOT HBM-bound (launch-latency-bound: one cluster launch or three kernels per hop):
        // a call that brings several hops runs their forward FFTs, multiply-accumulates and inverse FFTs as three launches
        // (spectra that stay in L2 between the hops of a batch: fused engines, and whatever else streams less than 48 MiB per hop;
        // an engine that streams more from HBM gains nothing from sharing launches -- config 5 -- and goes hop by hop)
        c->hb_max = (c->multi_hop && !c->mh_ok && (c->fused || tail_bytes < (uint64_t(48) << 20)) && (int) log2m <= single_cta_max_log2m(c) && g.P >= 1)
                        ? c->extra_slots + 1 : 1;
        // hops one pass carries at most: 8 (half units) where the prepared delay-line tiles stay small beside the IR unit
        c->mh_max = !c->mh_ok ? 1 : (!c->mh_packed ? 4 : (mh2_supported(g, 8) ? 8 : (mh2_supported(g, 4) ? 4 : 2)));
    }
    c->r_full = make_range(0, g.P, sms * per_sm);
    c->r_head = make_range(0, g.P ? 1 : 0, sms);
    c->r_tail = make_range(1, g.P ? g.P - 1 : 0, (sms - reserve) * per_sm);
    {
        // L2 policy of the streamed copies.  IR units: evict_first (read once per hop, gigabytes).  FDL tiles: evict_last while
        // the whole delay line (re-read every hop) can stay in the 126 MB L2; a delay line larger than that cannot be kept, and
        // marking it evict_last only costs bandwidth (config 5, 256 MiB: 95.0 -> 88.7 us per hop with evict_first).  Keeping a
        // SUBSET of the IR or of a large FDL resident with evict_last was measured too and is slower than plain streaming
        // (profiles/r1_l2_policy.txt).  HB_PIN_P / HB_PIN_S: experiments only.
        static const char *env_pp = getenv("HB_PIN_P"), *env_ps = getenv("HB_PIN_S");
        const uint64_t fdl_bytes = uint64_t(g.groups) * g.ins * g.n_bt * g.R * g.TBV * 16;
        Range *rs[3] = {&c->r_full, &c->r_head, &c->r_tail};
        for (Range *r : rs)
        {
            r->pin_p = env_pp ? (uint32_t) atoi(env_pp) : 0u;
            r->pin_s = env_ps ? (uint32_t) atoi(env_ps) : (fdl_bytes <= (uint64_t(64) << 20) ? g.R : 0u);
        }
    }
    g.G = c->r_full.G;
}

size_t h_vectors(const hb_conv *c)
{
    // groups * (n_ot*OT) padded rows * ins * max_length bins, in 16-byte vectors
    const uint32_t ot = choose_ot(c->outs);
    const uint32_t rows = ((c->outs + ot - 1) / ot) * ot;
    const size_t cpv = c->dtype == HB_F64 ? 1 : 2;
    return size_t(c->groups) * rows * c->ins * c->max_length / cpv;
}

void free_device(hb_conv *c)
{
    for (hb_conv::Reveal &r : c->reveals) cudaFree(r.side);
    c->reveals.clear();
    cudaFree(c->d_H); cudaFree(c->d_X); cudaFree(c->d_Hnyq); cudaFree(c->d_Xnyq); cudaFree(c->d_tw);
    c->d_H = c->d_X = c->d_Hnyq = c->d_Xnyq = c->d_tw = nullptr;
    c->d_S.release();
    c->d_St[0].release(); c->d_St[1].release(); c->d_trace.release();
    c->big.release(); c->d_nyq.release(); c->d_Smh.release(); c->d_part.release(); c->d_chain.release();
    if (c->s_tail) cudaStreamDestroy(c->s_tail);
    if (c->s_tail_b) cudaStreamDestroy(c->s_tail_b);
    c->s_tail_b = nullptr;
    if (c->ev_fwd) cudaEventDestroy(c->ev_fwd);
    for (int k = 0; k < 2; k++) if (c->ev_tail[k]) cudaEventDestroy(c->ev_tail[k]);
    c->s_tail = nullptr; c->ev_fwd = nullptr; c->ev_tail[0] = c->ev_tail[1] = nullptr;
    c->tail_valid = false;
    for (int k = 0; k < 2; k++) { c->d_xin[k].release(); c->d_yout[k].release(); }
    c->d_io_in.release(); c->d_io_out.release(); c->d_ir.release();
    c->h_in.release(); c->h_out.release(); c->h_ir.release();
    for (uint32_t r = 0; r < c->shard_world; r++)
        if (c->peers_attached && !c->peers_local && r != c->shard_rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    cudaFree(c->d_inbox);
    c->d_inbox = nullptr; c->peers_attached = false; c->peers_local = false; c->shard_world = 0;
    for (int k = 0; k < 2; k++)
    {
        c->h_blk[k].release(); c->h_inq[k].release(); c->d_inq[k].release(); c->d_blk[k].release();
        if (c->ev_blk[k]) cudaEventDestroy(c->ev_blk[k]);
        if (c->ev_inq[k]) cudaEventDestroy(c->ev_inq[k]);
        if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
        c->ev_blk[k] = c->ev_inq[k] = c->ev_done[k] = nullptr;
        c->blk_pending[k] = c->inq_pending[k] = c->done_pending[k] = false;
    }
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    c->s_h2d = c->s_d2h = nullptr;
    c->blk_valid = false;
}

// Nyquist side arrays are sized for the smallest hop the object may be switched to
size_t nyq_capacity(const hb_conv *c)
{
    const size_t minB = size_t(1) << (MIN_FFT_LOG2 - 1);
    return c->max_length / minB + c->extra_slots;
}

// allocate everything whose size depends on max_length (ctor and resize)
int alloc_capacity(hb_conv *c)
{
    for (hb_conv::Reveal &r : c->reveals) cudaFree(r.side);     // (resize puts them in place first)
    c->reveals.clear();
    cudaFree(c->d_H); cudaFree(c->d_X); cudaFree(c->d_Hnyq); cudaFree(c->d_Xnyq);
    c->d_H = c->d_X = c->d_Hnyq = c->d_Xnyq = nullptr;
    const size_t hbytes = h_vectors(c) * 16;
    const size_t maxB_ = (size_t(1) << c->max_fft_log2) >> 1;
    // look-ahead slots: 7 (multi-hop reuse on HBM-bound engines); 63 while they cost little memory, so that a launch-latency-bound
    // engine can take up to 64 hops of a call in one set of launches
    c->extra_slots = size_t(c->groups) * c->ins * 63 * maxB_ * 2 * c->esize() <= (size_t(32) << 20) ? 63u : MH_EXTRA;
    const size_t xbytes = size_t(c->groups) * c->ins * (c->max_length + c->extra_slots * maxB_) * 2 * c->esize();
    const size_t pmax = nyq_capacity(c);
    if (cudaMalloc(&c->d_H, std::max<size_t>(hbytes, 16)) != cudaSuccess ||
        cudaMalloc(&c->d_X, std::max<size_t>(xbytes, 16)) != cudaSuccess ||
        cudaMalloc(&c->d_Hnyq, std::max<size_t>(c->pairs() * pmax * c->esize(), 16)) != cudaSuccess ||
        cudaMalloc(&c->d_Xnyq, std::max<size_t>(size_t(c->groups) * c->ins * pmax * c->esize(), 16)) != cudaSuccess)
    {
        set_error("device allocation failed for %zu taps per pair (%s)", (size_t) c->max_length, cudaGetErrorString(cudaGetLastError()));
        return ERR_MEM_UNAVAILABLE;
    }
    // padded rows and never-set pairs must read as silence
    HB_CUDA(cudaMemsetAsync(c->d_H, 0, std::max<size_t>(hbytes, 16), c->stream));
    HB_CUDA(cudaMemsetAsync(c->d_Hnyq, 0, std::max<size_t>(c->pairs() * pmax * c->esize(), 16), c->stream));
    std::fill(c->nparts.begin(), c->nparts.end(), 0u);
    c->P = 0;
    c->need_reset = true;
    return ERR_NONE;
}

// opt in to > 48 KB of dynamic shared memory once per (kernel, device, size).  Every kernel of a hop also asks
// for the largest shared-memory carveout: in the overlapped schedule FFT CTAs are placed beside a resident
// multiply-accumulate CTA, and an SM keeps the carveout of the kernel that occupied it first -- sized for that
// kernel alone it would leave no room for the second one.
template <class K> int allow_smem(K kernel, size_t bytes)
{
    static std::mutex m;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    auto it = done.find(std::make_pair((const void *) kernel, dev));
    if (it == done.end())
    {
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared));
        it = done.insert(std::make_pair(std::make_pair((const void *) kernel, dev), size_t(48 * 1024))).first;
    }
    if (it->second >= bytes) return HB_OK;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    it->second = bytes;
    return HB_OK;
}

template <class T> size_t fft_smem(uint32_t log2m) { return size_t(padded_elems<HB_PADSH>(1u << log2m)) * sizeof(Cx<T>); }
// shared-memory copy of the twiddles of a real transform of 2^(log2m+1) points: 2^log2m entries behind the data
template <class T> size_t tw_smem(uint32_t log2m) { return (size_t(1) << log2m) * sizeof(Cx<T>); }
template <class T> int tw_fits(uint32_t log2m) { return fft_smem<T>(log2m) + tw_smem<T>(log2m) <= 200 * 1024 ? 1 : 0; }

// ---- kernel dispatch ------------------------------------------------------------------------------
template <class T, int XA, int OB>
int launch_cmac_inst(hb_conv *c, const Range &r, void *S, int variant, cudaStream_t st, uint32_t nb = 1, uint64_t set_stride = 0)
{
    typedef typename VecOf<T>::type V;
    const Geom &g = c->g;
    if (variant == 1)
    {
        // two tail streams: more than half an SM's shared memory per CTA, so that a CTA of the next tail is placed when one of the
        // running tail leaves and never beside it (two resident tails would leave no room for the FFT kernels of the critical path)
        const size_t smem = (c->tail_streams == 2 && r.kind == 2 && c->split) ? std::max<size_t>(c->cmac_smem, 116 * 1024) : c->cmac_smem;
        int rc = allow_smem(k_cmac_tma<T, XA, OB>, smem);
        if (rc) return rc;
        k_cmac_tma<T, XA, OB><<<r.G, 256, smem, st>>>(g, r, (const V *) c->d_H, (const V *) c->d_X, (V *) S, c->nstages);
    }
    else
    {
        int rc = allow_smem(k_cmac_ldg<T, XA, OB>, 0);
        if (rc) return rc;
        k_cmac_ldg<T, XA, OB><<<dim3(r.G, nb), 256, 0, st>>>(g, r, (const V *) c->d_H, (const V *) c->d_X, (V *) S, set_stride);
    }
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// one multiply-accumulate launch over the partitions of `r` into the partial segments S
// nb > 1 (direct-load variant only): the same partitions for nb consecutive hops in one launch, hop j against the frame j slots
// below r.slot, partial segments set_stride vectors apart
template <class T>
int launch_cmac(hb_conv *c, const Range &r, void *S, int variant, cudaStream_t st, uint32_t nb = 1, uint64_t set_stride = 0)
{
    if (!r.U) return HB_OK;
    const uint32_t key = c->g.XA * 16 + c->g.OB;
    switch (key)
    {
        case 1 * 16 + 1: return launch_cmac_inst<T, 1, 1>(c, r, S, variant, st, nb, set_stride);
        case 1 * 16 + 2: return launch_cmac_inst<T, 1, 2>(c, r, S, variant, st, nb, set_stride);
        case 1 * 16 + 4: return launch_cmac_inst<T, 1, 4>(c, r, S, variant, st, nb, set_stride);
        case 1 * 16 + 8: return launch_cmac_inst<T, 1, 8>(c, r, S, variant, st, nb, set_stride);
        case 2 * 16 + 1: return launch_cmac_inst<T, 2, 1>(c, r, S, variant, st, nb, set_stride);
        case 2 * 16 + 2: return launch_cmac_inst<T, 2, 2>(c, r, S, variant, st, nb, set_stride);
        case 2 * 16 + 4: return launch_cmac_inst<T, 2, 4>(c, r, S, variant, st, nb, set_stride);
        case 4 * 16 + 1: return launch_cmac_inst<T, 4, 1>(c, r, S, variant, st, nb, set_stride);
        case 4 * 16 + 2: return launch_cmac_inst<T, 4, 2>(c, r, S, variant, st, nb, set_stride);
        case 8 * 16 + 1: return launch_cmac_inst<T, 8, 1>(c, r, S, variant, st, nb, set_stride);
    }
    set_error("internal: no multiply-accumulate kernel for XA=%u OB=%u", c->g.XA, c->g.OB);
    return HB_ERR_UNSUPPORTED;
}

template <class T, int XA, int OB, int NH>
int launch_cmac_mh_inst(hb_conv *c, const Range &r, void *S, uint64_t set_stride, cudaStream_t st)
{
    typedef typename VecOf<T>::type V;
    const Geom &g = c->g;
    const size_t smem = size_t(c->mh_stages) * size_t(g.Q + NH * g.TBV) * 16 + size_t(c->mh_stages) * 8;
    int rc = allow_smem(k_cmac_tma_mh<T, XA, OB, NH>, smem);
    if (rc) return rc;
    k_cmac_tma_mh<T, XA, OB, NH><<<r.G, 288, smem, st>>>(g, r, (const V *) c->d_H, (const V *) c->d_X, (V *) S, c->mh_stages, set_stride);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// rows of an IR unit one CTA of a multi-hop launch works on: all of them, or half (8 hops: hb_conv_mh.cuh)
inline uint32_t mh_split(const hb_conv *c, int nh) { return c->mh_packed && nh == 8 ? 2u : 1u; }

// NH (2, 4 or 8) hops over one pass of the IR spectra; S holds NH partial-segment sets set_stride vectors apart
template <class T>
int launch_cmac_mh(hb_conv *c, const Range &r, void *S, int nh, uint64_t set_stride, cudaStream_t st)
{
    if (c->mh_packed) return launch_cmac_mh2(c->g, r, c->d_H, c->d_X, S, nh, set_stride, st);
    const uint32_t key = (c->g.XA * 16 + c->g.OB) * 8 + (uint32_t) nh;
    switch (key)
    {
        case (1 * 16 + 8) * 8 + 2: return launch_cmac_mh_inst<T, 1, 8, 2>(c, r, S, set_stride, st);
        case (1 * 16 + 8) * 8 + 4: return launch_cmac_mh_inst<T, 1, 8, 4>(c, r, S, set_stride, st);
        case (2 * 16 + 4) * 8 + 2: return launch_cmac_mh_inst<T, 2, 4, 2>(c, r, S, set_stride, st);
        case (2 * 16 + 4) * 8 + 4: return launch_cmac_mh_inst<T, 2, 4, 4>(c, r, S, set_stride, st);
    }
    set_error("internal: no multi-hop kernel for XA=%u OB=%u NH=%d", c->g.XA, c->g.OB, nh);
    return HB_ERR_UNSUPPORTED;
}

template <class T, int EPT>
int launch_fwd_ept(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, cudaStream_t st, uint32_t nh)
{
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const int stage_tw = tw_fits<T>(log2m);
    const size_t smem = fft_smem<T>(log2m) + (stage_tw ? tw_smem<T>(log2m) : 0);
    int rc = allow_smem(k_fwd<T, EPT>, smem);
    if (rc) return rc;
    k_fwd<T, EPT><<<dim3(g.groups * g.ins, nh), fft_threads(log2m, EPT), smem, st>>>(g, prev, prev_ld, newest, new_ld, save, save_ld, (Cx<T> *) c->d_X, (T *) c->d_Xnyq,
                                                                        (const Cx<T> *) c->d_tw, c->tw_log2, stage_tw);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T> bool is_big(const hb_conv *c) { return (int) c->g.log2n - 1 > single_cta_max_log2m(c); }
inline dim3 bigc_grid(uint32_t B, size_t batch) { return dim3((unsigned) std::min<size_t>((B + 255) / 256, 256), (unsigned) batch); }

// forward transform of every input channel at sizes above the single-CTA limit (hb_conv_big.cuh)
template <class T>
int launch_fwd_big(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, cudaStream_t st)
{
    const Geom &g = c->g;
    const int m = (int) g.log2n - 1;
    const size_t rows = size_t(g.groups) * g.ins;
    int rc;
    if ((rc = c->big.ensure<T>(m, std::max(rows, size_t(g.groups) * g.outs)))) return rc;
    Cx<T> *z1 = (Cx<T> *) c->big.z1.p, *z2 = (Cx<T> *) c->big.z2.p;
    const Cx<T> *tw = (const Cx<T> *) c->d_tw;
    k_bigc_pack<T><<<bigc_grid(g.B, rows), 256, 0, st>>>(g, prev, prev_ld, newest, new_ld, save, save_ld, z1);
    HB_LAUNCH_CHECK();
    if ((rc = big_cfft<T>(z1, z2, z1, m, rows, tw, c->tw_log2, st))) return rc;
    if ((rc = big_split<T>(z1, m, 0, rows, tw, c->tw_log2, st))) return rc;
    k_bigc_to_fdl<T><<<bigc_grid(g.B, rows), 256, 0, st>>>(g, z1, (Cx<T> *) c->d_X, (T *) c->d_Xnyq);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// forward transform of every input channel on clusters of 8 CTAs (hb_conv_cluster.cuh)
template <class T>
int launch_fwd_cl(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, cudaStream_t st)
{
    const Geom &g = c->g;
    const int m = (int) g.log2n - 1;
    const size_t smem = cl_fwd_smem<T>(m);
    int rc = allow_smem(k_fwd_cl<T>, smem);
    if (rc) return rc;
    k_fwd_cl<T><<<g.groups * g.ins * CL_CS, (1u << m) / (CL_CS * CL_EPT), smem, st>>>(g, prev, prev_ld, newest, new_ld, save, save_ld, (Cx<T> *) c->d_X, (T *) c->d_Xnyq,
                                                                                    (const Cx<T> *) c->d_tw, c->tw_log2);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// nh > 1: the forward transforms of nh consecutive hops in one launch (hop j from block j of the caller's rows into slot
// c->g.slot - j; single-CTA transforms only)
template <class T>
int launch_fwd(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, cudaStream_t st, uint32_t nh = 1)
{
    if (is_big<T>(c)) return launch_fwd_big<T>(c, prev, prev_ld, newest, new_ld, save, save_ld, st);
    if (nh == 1 && use_cluster_fft(c)) return launch_fwd_cl<T>(c, prev, prev_ld, newest, new_ld, save, save_ld, st);
    HB_EPT_DISPATCH(c->g.log2n - 1, return launch_fwd_ept<T, EPT>(c, prev, prev_ld, newest, new_ld, save, save_ld, st, nh));
}

// what k_inv does besides the transform: where the block goes and which finished block it hands over first
template <class T> struct InvIO
{
    T *yout; size_t ld, off; int add_result;
    const T *carry_src; size_t carry_src_ld; T *carry_dst; size_t carry_dst_ld; int add_carry;
};

template <class T, int EPT>
int launch_inv_ept(hb_conv *c, const SegSets &sets, const InvIO<T> &io, cudaStream_t st, const PeerOut &peer, uint32_t nh, const InvBatch &ib)
{
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const int stage_tw = tw_fits<T>(log2m);
    const size_t smem = fft_smem<T>(log2m) + (stage_tw ? tw_smem<T>(log2m) : 0);
    int rc = allow_smem(k_inv<T, EPT>, smem);
    if (rc) return rc;
    k_inv<T, EPT><<<dim3(g.groups * g.outs, nh), fft_threads(log2m, EPT), smem, st>>>(g, sets, (const T *) c->d_Xnyq, (const T *) c->d_Hnyq, io.yout, io.ld, io.off,
                                                                         io.add_result, io.carry_src, io.carry_src_ld, io.carry_dst, io.carry_dst_ld, io.add_carry,
                                                                         (const Cx<T> *) c->d_tw, c->tw_log2, stage_tw, peer, ib);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// inverse transform of every output channel at sizes above the single-CTA limit (hb_conv_big.cuh)
template <class T>
int launch_inv_big(hb_conv *c, const SegSets &sets, const InvIO<T> &io, cudaStream_t st)
{
    const Geom &g = c->g;
    const int m = (int) g.log2n - 1;
    const size_t rows = size_t(g.groups) * g.outs;
    int rc;
    if ((rc = c->big.ensure<T>(m, std::max(rows, size_t(g.groups) * g.ins))) || (rc = c->d_nyq.ensure(rows * sizeof(T)))) return rc;
    Cx<T> *z1 = (Cx<T> *) c->big.z1.p, *z2 = (Cx<T> *) c->big.z2.p;
    const Cx<T> *tw = (const Cx<T> *) c->d_tw;
    k_bigc_nyq<T><<<(unsigned) rows, 256, 0, st>>>(g, (const T *) c->d_Xnyq, (const T *) c->d_Hnyq, (T *) c->d_nyq.p);
    HB_LAUNCH_CHECK();
    k_bigc_gather<T><<<bigc_grid(g.B, rows), 256, 0, st>>>(g, sets, (const T *) c->d_nyq.p, z1, io.carry_src, io.carry_src_ld, io.carry_dst, io.carry_dst_ld, io.add_carry);
    HB_LAUNCH_CHECK();
    if ((rc = big_split<T>(z1, m, 1, rows, tw, c->tw_log2, st))) return rc;
    k_big_exchange<T><<<(unsigned) std::min<size_t>((rows * g.B + 255) / 256, 2048), 256, 0, st>>>(z1, rows * g.B);
    HB_LAUNCH_CHECK();
    if ((rc = big_cfft<T>(z1, z2, z1, m, rows, tw, c->tw_log2, st))) return rc;
    k_bigc_store<T><<<bigc_grid(g.B / 2, rows), 256, 0, st>>>(g, z1, io.yout, io.ld, io.off, io.add_result);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// inverse transform of every output channel on clusters of 8 CTAs (hb_conv_cluster.cuh)
template <class T>
int launch_inv_cl(hb_conv *c, const SegSets &sets, const InvIO<T> &io, cudaStream_t st)
{
    const Geom &g = c->g;
    const int m = (int) g.log2n - 1;
    const size_t smem = cl_inv_smem<T>(m, g.n_bt);
    int rc = allow_smem(k_inv_cl<T>, smem);
    if (rc) return rc;
    k_inv_cl<T><<<g.groups * g.outs * CL_CS, (1u << m) / (CL_CS * CL_EPT), smem, st>>>(g, sets, (const T *) c->d_Xnyq, (const T *) c->d_Hnyq, io.yout, io.ld, io.off, io.add_result,
                                                                                     io.carry_src, io.carry_src_ld, io.carry_dst, io.carry_dst_ld, io.add_carry,
                                                                                     (const Cx<T> *) c->d_tw, c->tw_log2);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

inline InvBatch no_batch() { InvBatch ib; memset(&ib, 0, sizeof(ib)); ib.last_j = -1; return ib; }

// nh > 1: the inverse transforms of nh consecutive hops in one launch (hop j: segment sets ib.set_stride apart, block j of the
// caller's row; see InvBatch)
template <class T>
int launch_inv(hb_conv *c, const SegSets &sets, const InvIO<T> &io, cudaStream_t st, const PeerOut &peer = PeerOut(), uint32_t nh = 1,
               const InvBatch &ib = no_batch())
{
    if (is_big<T>(c))
    {
        if (!peer.world) return launch_inv_big<T>(c, sets, io, st);
        // fused multi-GPU exchange: the chain leaves the partial blocks of all outputs in a local buffer, k_shard_deliver hands them on
        const Geom &g = c->g;
        const size_t rows = size_t(g.groups) * g.outs;
        int rc = c->d_part.ensure(rows * g.B * sizeof(T));
        if (rc) return rc;
        const InvIO<T> local = {(T *) c->d_part.p, g.B, 0, 0, nullptr, 0, nullptr, 0, 0};
        if ((rc = launch_inv_big<T>(c, sets, local, st))) return rc;
        k_shard_deliver<T><<<(unsigned) rows, 256, 0, st>>>((const T *) c->d_part.p, g.B, peer, g.B);
        HB_LAUNCH_CHECK();
        return HB_OK;
    }
    if (nh == 1 && ib.last_j < 0 && !peer.world && use_cluster_fft(c)) return launch_inv_cl<T>(c, sets, io, st);
    HB_EPT_DISPATCH(c->g.log2n - 1, return launch_inv_ept<T, EPT>(c, sets, io, st, peer, nh, ib));
}

// the whole hop in one cluster launch (hb_conv_fused.cuh); eligibility is decided in plan_geometry
template <class T, int MAXT>
int launch_fused_t(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, const InvIO<T> &io, cudaStream_t st)
{
    const Geom &g = c->g;
    constexpr int EPT = 8;
    const uint32_t B = g.B, cs = c->fused_cs;
    const size_t smem = (size_t(padded_elems<HB_PADSH>(B)) + 2 * size_t(B)) * sizeof(Cx<T>);
    auto kernel = k_hop_fused<T, EPT, MAXT>;
    int rc = allow_smem(kernel, smem);
    if (rc) return rc;
    if (cs > 8)
    {
        static std::mutex m;
        static std::map<std::pair<const void *, int>, bool> done;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(m);
        bool &ok = done[std::make_pair((const void *) kernel, dev)];
        if (!ok) { HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); ok = true; }
    }
    FusedArgs fa;
    fa.cs = cs;
    fa.tail_items = g.P > 2 ? g.ins * (g.P - 2) : 0;
    // chained: this hop follows a fused hop of this engine directly (nothing else was enqueued in between by the engine, and
    // nothing by the caller if the stream is the caller's -- which is what hop_overlap = 2 declares) and may run beside it
    const bool overlap_allowed = c->hop_overlap == 2 || (c->hop_overlap == 1 && st == c->stream);
    fa.chained = (overlap_allowed && c->chain_ok && c->chain_stream == st && c->chain_n > 0) ? 1u : 0u;
    // a call fed with the output rows of one of the last calls (the engine in a feedback loop) relies on stream order for its input
    {
        const size_t rows_in = size_t(g.groups) * g.ins, rows_out = size_t(g.groups) * g.outs;
        const char *ib = (const char *) newest, *ie = (const char *) (newest + (rows_in - 1) * new_ld + B);
        for (const hb_conv::ByteRange &r : c->chain_out)
            if (r.b && ib < r.e && r.b < ie) fa.chained = 0;
        auto note = [&](const T *p, size_t ld)
        {
            hb_conv::ByteRange &r = c->chain_out[c->chain_out_pos++ % 16];
            r.b = (const char *) p;
            r.e = p ? (const char *) (p + (rows_out - 1) * ld + B) : nullptr;
        };
        note(io.carry_dst, io.carry_dst_ld);
        note(io.yout ? io.yout + io.off : nullptr, io.ld);
    }
    // only hops that a chained hop may follow take part in the count (the fences of the counter updates cost a strict hop about a
    // microsecond): chain_n numbers those, and a chained hop always follows one of them directly
    fa.bump = overlap_allowed ? 1u : 0u;
    fa.tail_early = c->last_hop_chained ? 0u : 1u;
    fa.writers = g.groups * std::min<uint32_t>(cs, g.ins);
    fa.clusters = g.groups * g.outs;
    fa.depth = std::min<uint32_t>(g.R - g.P, 8u);
    fa.n = overlap_allowed ? ++c->chain_n : 0;
    fa.sync = (unsigned long long *) c->d_chain.p;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(g.groups * g.outs * cs);
    cfg.blockDim = dim3(std::max<uint32_t>(std::min<uint32_t>(B / EPT, 512u), 128u));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    // programmatic dependent launch: this hop may start beside the previous kernel of the stream and runs what does not depend on
    // it (twiddles, newest block, partitions >= 2) until griddepcontrol.wait (hb_conv_fused.cuh); HB_NO_PDL: experiments only
    static const char *env_pdl = getenv("HB_NO_PDL");
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = (env_pdl && atoi(env_pdl)) ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    HB_CUDA(cudaLaunchKernelEx(&cfg, kernel, g, fa, prev, prev_ld, newest, new_ld, save, save_ld, (const Cx<T> *) c->d_H, (Cx<T> *) c->d_X,
                               (const T *) c->d_Hnyq, (T *) c->d_Xnyq, io.yout, io.ld, io.off, io.add_result,
                               io.carry_src, io.carry_src_ld, io.carry_dst, io.carry_dst_ld, io.add_carry, (const Cx<T> *) c->d_tw, c->tw_log2));
    HB_LAUNCH_CHECK();
    c->chain_ok = overlap_allowed;
    c->chain_stream = st;
    c->last_hop_chained = fa.chained != 0;
    return HB_OK;
}

template <class T>
int launch_fused(hb_conv *c, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, const InvIO<T> &io, cudaStream_t st)
{
    return c->g.B / 8 <= 256 ? launch_fused_t<T, 256>(c, prev, prev_ld, newest, new_ld, save, save_ld, io, st)
                             : launch_fused_t<T, 512>(c, prev, prev_ld, newest, new_ld, save, save_ld, io, st);
}

template <class T, int EPT>
int launch_ir_ept(hb_conv *c, const T *d_ir, size_t taps, uint32_t grp, uint32_t in, uint32_t o, uint32_t nwrite, cudaStream_t st, void *side)
{
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const size_t smem = fft_smem<T>(log2m);
    int rc = allow_smem(k_ir<T, EPT>, smem);
    if (rc) return rc;
    k_ir<T, EPT><<<nwrite, fft_threads(log2m, EPT), smem, st>>>(g, d_ir, taps, grp, in, o, (Cx<T> *) c->d_H, (T *) c->d_Hnyq, (const Cx<T> *) c->d_tw, c->tw_log2,
                                                                (Cx<T> *) side, nwrite);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int launch_ir_big(hb_conv *c, const T *d_ir, size_t taps, uint32_t grp, uint32_t in, uint32_t o, uint32_t nwrite, cudaStream_t st)
{
    const Geom &g = c->g;
    const int m = (int) g.log2n - 1;
    const Cx<T> *tw = (const Cx<T> *) c->d_tw;
    // partitions in batches that keep the scratch modest (the engine's own scratch is sized for its channels)
    const uint32_t batch_max = (uint32_t) std::max<size_t>(1, (size_t(64) << 20) / (size_t(g.B) * sizeof(Cx<T>)));
    BigScratch scratch;
    int rc = HB_OK;
    if ((rc = scratch.ensure<T>(m, std::min(nwrite, batch_max)))) return rc;
    Cx<T> *z1 = (Cx<T> *) scratch.z1.p, *z2 = (Cx<T> *) scratch.z2.p;
    for (uint32_t p0 = 0; p0 < nwrite && rc == HB_OK; p0 += batch_max)
    {
        const uint32_t nb = std::min(batch_max, nwrite - p0);
        Geom gp = g;                                     // kernels index partitions from blockIdx.y: shift the base instead
        const size_t skip = size_t(p0) * g.B;
        const T *src = d_ir + std::min(skip, taps);
        const size_t left = taps > skip ? taps - skip : 0;
        k_bigc_ir_pack<T><<<bigc_grid(g.B, nb), 256, 0, st>>>(gp, src, left, z1);
        count_launch();
        if ((rc = big_cfft<T>(z1, z2, z1, m, nb, tw, c->tw_log2, st))) break;
        if ((rc = big_split<T>(z1, m, 0, nb, tw, c->tw_log2, st))) break;
        // unit addresses advance by Q per partition: partition p0 + k of this batch = partition k of a layout shifted by p0 units
        Cx<T> *Hs = (Cx<T> *) c->d_H + size_t(p0) * g.Q * VecOf<T>::CPV;
        T *Hn = (T *) c->d_Hnyq + p0;
        k_bigc_to_h<T><<<bigc_grid(g.B, nb), 256, 0, st>>>(gp, z1, grp, in, o, Hs, Hn);
        count_launch();
    }
    cudaError_t e = cudaStreamSynchronize(st);            // the scratch is freed on return
    scratch.release();
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("impulse-response transform -> %s", cudaGetErrorString(e)); return HB_ERR_CUDA; }
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("impulse-response kernels -> %s", cudaGetErrorString(e)); return HB_ERR_CUDA; }
    return HB_OK;
}

// side != nullptr (single-CTA transform sizes only): the spectra of the nwrite partitions go to that private buffer (k_ir)
template <class T>
int launch_ir(hb_conv *c, const T *d_ir, size_t taps, uint32_t grp, uint32_t in, uint32_t o, uint32_t nwrite, cudaStream_t st, void *side = nullptr)
{
    if (!nwrite) return HB_OK;
    if (is_big<T>(c)) return launch_ir_big<T>(c, d_ir, taps, grp, in, o, nwrite, st);
    HB_EPT_DISPATCH(c->g.log2n - 1, return launch_ir_ept<T, EPT>(c, d_ir, taps, grp, in, o, nwrite, st, side));
}

// ---- pairs restarted while the stream runs ----------------------------------------------------------
template <class T>
int pair_copy(hb_conv *c, const hb_conv::Reveal &r, uint32_t p0, uint32_t count, int mode, cudaStream_t st)
{
    if (!count) return HB_OK;
    k_pair_copy<T><<<count, 256, 0, st>>>(c->g, (Cx<T> *) r.side, r.np, r.grp, r.in, r.out, p0, (Cx<T> *) c->d_H, (T *) c->d_Hnyq, mode);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// before the multiply-accumulate of a hop: every restarted pair gets the partition whose frame is the first one recorded after
// its restart (partition p at age p), so that it never meets input from before the restart
template <class T>
int advance_reveals(hb_conv *c, cudaStream_t st)
{
    for (size_t k = 0; k < c->reveals.size();)
    {
        hb_conv::Reveal &r = c->reveals[k];
        const uint32_t want = std::min(r.np, r.age + 1);
        int rc = pair_copy<T>(c, r, r.shown, want - r.shown, 0, st);
        if (rc) return rc;
        r.shown = want;
        r.age++;
        if (r.shown == r.np)
        {
            // the buffer may still be read by the copy just enqueued: freed behind the stream
            HB_CUDA(cudaFreeAsync(r.side, st));
            c->reveals.erase(c->reveals.begin() + k);
        }
        else k++;
    }
    return HB_OK;
}

// everything in place at once (whole-engine reset, resize): after a reset the history is silence anyway
template <class T>
int finish_reveals(hb_conv *c, cudaStream_t st)
{
    for (hb_conv::Reveal &r : c->reveals)
    {
        int rc = pair_copy<T>(c, r, r.shown, r.np - r.shown, 0, st);
        if (rc) return rc;
        HB_CUDA(cudaFreeAsync(r.side, st));
    }
    c->reveals.clear();
    return HB_OK;
}

void drop_reveal(hb_conv *c, uint32_t grp, uint32_t in, uint32_t o)
{
    for (size_t k = 0; k < c->reveals.size(); k++)
        if (c->reveals[k].grp == grp && c->reveals[k].in == in && c->reveals[k].out == o)
        {
            cudaFree(c->reveals[k].side);
            c->reveals.erase(c->reveals.begin() + k);
            return;
        }
}

template <class T>
int launch_rows(T *dst, size_t dld, const T *src, size_t sld, size_t n, size_t rows, int add, cudaStream_t st)
{
    if (!n || !rows) return HB_OK;
    dim3 grid((unsigned) std::min<size_t>((n + 255) / 256, 64), (unsigned) rows);
    k_rows<T><<<grid, 256, 0, st>>>(dst, dld, src, sld, n, add);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

bool fft_supported(const hb_conv *c, uintptr_t log2n)
{
    (void) c;
    return (int) log2n - 1 <= BIG_MAX_LOG2;           // single CTA up to SmemFftLimit, four-step above (hb_conv_big.cuh)
}

// setFFTSize semantics (PartitionedConvolve.cpp:131-154)
int apply_fft_size(hb_conv *c, uintptr_t fft_size)
{
    const uintptr_t l2 = ceil_log2(fft_size);
    int error = ERR_NONE;
    if (l2 < MIN_FFT_LOG2 || l2 > c->max_fft_log2) return ERR_FFT_SIZE_OUT_OF_RANGE;
    if (fft_size != (uintptr_t(1) << l2)) error = ERR_FFT_SIZE_NON_POWER_OF_TWO;
    if (l2 != c->fft_log2)
    {
        // every loaded IR is dropped (PartitionedConvolve.cpp:145-149).  The spectra are tiled per FFT size and the
        // multiply-accumulate kernels have no per-pair mask, so what is left in the arrays from the old size must read as
        // silence for the pairs that are not set again (or set shorter than the longest pair)
        if (c->s_tail) HB_CUDA(cudaStreamSynchronize(c->s_tail));
        if (c->s_tail_b) HB_CUDA(cudaStreamSynchronize(c->s_tail_b));
        HB_CUDA(cudaMemsetAsync(c->d_H, 0, std::max<size_t>(h_vectors(c) * 16, 16), c->stream));
        HB_CUDA(cudaMemsetAsync(c->d_Hnyq, 0, std::max<size_t>(c->pairs() * nyq_capacity(c) * c->esize(), 16), c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        c->tail_valid = false;
        for (hb_conv::Reveal &r : c->reveals) cudaFree(r.side);
        c->reveals.clear();
        std::fill(c->nparts.begin(), c->nparts.end(), 0u);
        c->P = 0;
        c->fft_log2 = l2;
        c->need_reset = true;
    }
    return error;
}

// make staging rows hold at least `need_x` / `need_y` elements, keeping the retained heads
template <class T>
int ensure_staging(hb_conv *c, size_t need_x, size_t need_y, bool preserve, cudaStream_t st)
{
    const size_t B = c->g.B;
    if (need_x > c->xin_ld)
    {
        const size_t ld = ((need_x * 3 / 2) + 63) & ~size_t(63);
        const size_t rows = size_t(c->groups) * c->ins;
        DevBuf n0, n1;
        int rc;
        if ((rc = n0.ensure(rows * ld * sizeof(T))) || (rc = n1.ensure(rows * ld * sizeof(T)))) return rc;
        if (c->xin_ld && preserve)
        {
            rc = launch_rows<T>((T *) n0.p, ld, (const T *) c->d_xin[c->cur].p + c->x_tail, c->xin_ld, B + c->rw, rows, 0, st);
            if (rc) return rc;
            HB_CUDA(cudaStreamSynchronize(st));
        }
        else
            HB_CUDA(cudaMemsetAsync(n0.p, 0, rows * ld * sizeof(T), st));
        c->d_xin[0].release(); c->d_xin[1].release();
        c->d_xin[0] = n0; c->d_xin[1] = n1;
        if (c->cur == 1) std::swap(c->d_xin[0], c->d_xin[1]);
        c->xin_ld = ld;
        c->x_tail = 0;
    }
    if (need_y > c->yout_ld)
    {
        const size_t ld = ((need_y * 3 / 2) + 63) & ~size_t(63);
        const size_t rows = size_t(c->groups) * c->outs;
        DevBuf n0, n1;
        int rc;
        if ((rc = n0.ensure(rows * ld * sizeof(T))) || (rc = n1.ensure(rows * ld * sizeof(T)))) return rc;
        if (c->yout_ld && preserve)
        {
            rc = launch_rows<T>((T *) n0.p, ld, (const T *) c->d_yout[c->cur].p + c->y_tail, c->yout_ld, B, rows, 0, st);
            if (rc) return rc;
            HB_CUDA(cudaStreamSynchronize(st));
        }
        else
            HB_CUDA(cudaMemsetAsync(n0.p, 0, rows * ld * sizeof(T), st));
        c->d_yout[0].release(); c->d_yout[1].release();
        c->d_yout[0] = n0; c->d_yout[1] = n1;
        if (c->cur == 1) std::swap(c->d_yout[0], c->d_yout[1]);
        c->yout_ld = ld;
        c->y_tail = 0;
    }
    return HB_OK;
}

// the reset block of process (PartitionedConvolve.cpp:267-290)
template <class T>
int do_reset(hb_conv *c, cudaStream_t st)
{
    // a tail launched ahead for a hop that will not come any more may still be running on its stream
    if (c->s_tail) HB_CUDA(cudaStreamSynchronize(c->s_tail));
    if (c->s_tail_b) HB_CUDA(cudaStreamSynchronize(c->s_tail_b));
    c->tail_valid = false;
    c->tail_missing = false;
    plan_geometry(c);
    const Geom &g = c->g;
    int rc = c->d_S.ensure(std::max<size_t>((size_t(std::max(c->r_full.G, c->r_head.G)) + g.tiles) * g.Q * 16, 16));
    if (rc) return rc;
    if (c->split)
    {
        for (int k = 0; k < 2; k++)
            if ((rc = c->d_St[k].ensure(std::max<size_t>((size_t(c->r_tail.G) + g.tiles) * g.Q * 16, 16)))) return rc;
        if (!c->s_tail) HB_CUDA(cudaStreamCreateWithFlags(&c->s_tail, cudaStreamNonBlocking));
        if (!c->s_tail_b) HB_CUDA(cudaStreamCreateWithFlags(&c->s_tail_b, cudaStreamNonBlocking));
        if (!c->ev_fwd) HB_CUDA(cudaEventCreateWithFlags(&c->ev_fwd, cudaEventDisableTiming));
        for (int k = 0; k < 2; k++)
            if (!c->ev_tail[k]) HB_CUDA(cudaEventCreateWithFlags(&c->ev_tail[k], cudaEventDisableTiming));
    }
    if ((rc = finish_reveals<T>(c, st))) return rc;
    // FDL silence: stale slots are masked in the reference by mValidPartitions (:285,373); zeros do the same
    HB_CUDA(cudaMemsetAsync(c->d_X, 0, size_t(g.groups) * g.ins * g.R * g.B * 2 * sizeof(T), st));
    HB_CUDA(cudaMemsetAsync(c->d_Xnyq, 0, size_t(g.groups) * g.ins * g.R * sizeof(T), st));
    c->rw = c->reset_offset < 0 ? 0 : uintptr_t(c->reset_offset) % g.B;
    c->g.slot = 0;
    if ((rc = c->d_chain.ensure(64))) return rc;
    HB_CUDA(cudaMemsetAsync(c->d_chain.p, 0, 64, st));
    c->chain_n = 0;
    c->chain_ok = false;
    // staging rows start as silence: previous hop, pending samples and previous result block
    if (c->xin_ld) HB_CUDA(cudaMemsetAsync(c->d_xin[c->cur].p, 0, size_t(g.groups) * g.ins * c->xin_ld * sizeof(T), st));
    if (c->yout_ld) HB_CUDA(cudaMemsetAsync(c->d_yout[c->cur].p, 0, size_t(g.groups) * g.outs * c->yout_ld * sizeof(T), st));
    c->x_tail = c->y_tail = 0;
    c->need_reset = false;
    c->blk_valid = false;
    return HB_OK;
}

// the first thing process does after set / reset / a size change (PartitionedConvolve.cpp:267-290)
template <class T>
int ensure_ready(hb_conv *c, size_t n, cudaStream_t st)
{
    if (!c->need_reset) return HB_OK;
    int rc;
    plan_geometry(c);
    if ((rc = ensure_staging<T>(c, 2 * size_t(c->g.B) + n, 2 * size_t(c->g.B) + n, false, st))) return rc;
    return do_reset<T>(c, st);
}

// fold the recorded event intervals into prof_ms (synchronises on the last recorded events)
constexpr size_t PROF_EV = 7;      // per hop: 5 marks on the launching stream, 2 on the tail stream
int drain_profile(hb_conv *c)
{
    if (!c->ev_used) return HB_OK;
    HB_CUDA(cudaEventSynchronize(c->ev[(c->ev_used - 1) * PROF_EV + 4]));
    if (c->ev_has_tail[c->ev_used - 1]) HB_CUDA(cudaEventSynchronize(c->ev[(c->ev_used - 1) * PROF_EV + 6]));
    for (size_t h = 0; h < c->ev_used; h++)
    {
        const cudaEvent_t *e = &c->ev[h * PROF_EV];
        for (int j = 0; j < 4; j++)
        {
            float ms = 0.f;
            HB_CUDA(cudaEventElapsedTime(&ms, e[j], e[j + 1]));
            c->prof_ms[j] += ms;
        }
        if (c->ev_has_tail[h])
        {
            float ms = 0.f;
            HB_CUDA(cudaEventElapsedTime(&ms, e[5], e[6]));
            c->prof_ms[4] += ms;
        }
        c->prof_hops++;
    }
    c->ev_used = 0;
    return HB_OK;
}

// events of the hop being recorded (allocates the next block of PROF_EV events)
int profile_begin_hop(hb_conv *c, cudaEvent_t **out)
{
    if (c->ev_used * PROF_EV == c->ev.size())
    {
        if (c->ev_used >= 256) { int rc = drain_profile(c); if (rc) return rc; }
        else
        {
            for (size_t k = 0; k < PROF_EV; k++)
            {
                cudaEvent_t e;
                HB_CUDA(cudaEventCreate(&e));
                c->ev.push_back(e);
            }
            c->ev_has_tail.push_back(0);
        }
    }
    c->ev_has_tail[c->ev_used] = 0;
    *out = &c->ev[c->ev_used * PROF_EV];
    c->ev_used++;
    return HB_OK;
}

// Overlapped schedule: launch the tail of the NEXT hop on the tail stream.  It needs nothing newer than the spectrum at
// c->g.slot (just written on st): its frame will go to slot - 1, so partition p meets slot - 1 + p -- the newest spectrum
// at p = 1, the oldest one kept at p = P - 1.  Leaves tail_par / tail_valid describing that launch.
template <class T>
int launch_tail_ahead(hb_conv *c, cudaStream_t st, cudaEvent_t *pe)
{
    int r;
    HB_CUDA(cudaEventRecord(c->ev_fwd, st));
    const int np = c->tail_par ^ 1;
    Range rt = c->r_tail;
    rt.slot = c->g.slot ? c->g.slot - 1 : c->g.R - 1;
    rt.kind = 2;
    // (two tail streams: this launch waits for the forward FFTs only, not for the tail that is still running on the other
    // stream; its segment set np was last read by the inverse FFTs two hops back, which precede ev_fwd on st)
    cudaStream_t ts = (c->tail_streams == 2 && np) ? c->s_tail_b : c->s_tail;
    HB_CUDA(cudaStreamWaitEvent(ts, c->ev_fwd, 0));
    if (pe) { HB_CUDA(cudaEventRecord(pe[5], ts)); c->ev_has_tail[c->ev_used - 1] = 1; }
    if ((r = launch_cmac<T>(c, rt, c->d_St[np].p, c->variant, ts))) return r;
    if (pe) HB_CUDA(cudaEventRecord(pe[6], ts));
    HB_CUDA(cudaEventRecord(c->ev_tail[np], ts));
    c->tail_par = np;
    c->tail_valid = true;
    return HB_OK;
}

// One hop: forward FFTs of the newest frame, multiply-accumulate, inverse FFTs (PartitionedConvolve.cpp:352-377).
// Serial schedule: the three kernels in a row on st.  Overlapped schedule: see hb_conv::schedule.
template <class T>
int launch_hop(hb_conv *c, cudaStream_t st, const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld,
               const InvIO<T> &io, const PeerOut &peer)
{
    int r;
    cudaEvent_t *pe = nullptr;
    if (c->profiling && (r = profile_begin_hop(c, &pe))) return r;
    // newest spectrum goes one slot below the previous one (mInputPosition--, cpp:374)
    c->g.slot = c->g.slot ? c->g.slot - 1 : c->g.R - 1;
    c->g.hop++;
    c->g.trace = (unsigned long long *) c->d_trace.p;
    // restarted pairs: the partition that becomes valid at this hop is put in place first; while any is pending the hop runs all its
    // partitions itself (a tail computed a hop ahead would need the next partition already)
    const bool revealing = !c->reveals.empty();
    if (revealing && (r = advance_reveals<T>(c, st))) return r;
    if (pe) HB_CUDA(cudaEventRecord(pe[0], st));
    if (revealing || pe || !c->fused || peer.world) c->chain_ok = false;      // something else sits between this hop and the last fused one
    if (c->fused && !peer.world)
    {
        if ((r = launch_fused<T>(c, prev, prev_ld, newest, new_ld, save, save_ld, io, st))) return r;
        if (pe) for (int k = 1; k <= 4; k++) HB_CUDA(cudaEventRecord(pe[k], st));
        return HB_OK;
    }
    if ((r = launch_fwd<T>(c, prev, prev_ld, newest, new_ld, save, save_ld, st))) return r;
    if (pe) HB_CUDA(cudaEventRecord(pe[1], st));
    SegSets sets;
    memset(&sets, 0, sizeof(sets));
    if (!c->split || revealing)
    {
        Range rf = c->r_full;
        rf.slot = c->g.slot; rf.kind = 2;
        if ((r = launch_cmac<T>(c, rf, c->d_S.p, c->variant, st))) return r;
        if (pe) { HB_CUDA(cudaEventRecord(pe[2], st)); HB_CUDA(cudaEventRecord(pe[3], st)); }
        sets.n = 1;
        sets.s[0].S = c->d_S.p; sets.s[0].U = rf.U; sets.s[0].upt = rf.upt; sets.s[0].G = rf.G;
        if (c->split) { c->tail_valid = false; c->tail_missing = true; }     // the overlapped schedule resumes with a full hop
    }
    else
    {
        // fork: the tail of the NEXT hop (launch_tail_ahead)
        const bool had_tail = c->tail_valid;
        const int old_par = c->tail_par;
        if ((r = launch_tail_ahead<T>(c, st, pe))) return r;
        if (c->tail_missing)
        {
            // the hop after a multi-hop batch: its tail was not computed ahead, so it runs all partitions itself
            c->tail_missing = false;
            Range rf = c->r_full;
            rf.slot = c->g.slot; rf.kind = 2;
            if ((r = launch_cmac<T>(c, rf, c->d_S.p, c->variant, st))) return r;
            if (pe) { HB_CUDA(cudaEventRecord(pe[2], st)); HB_CUDA(cudaEventRecord(pe[3], st)); }
            sets.n = 1;
            sets.s[0].S = c->d_S.p; sets.s[0].U = rf.U; sets.s[0].upt = rf.upt; sets.s[0].G = rf.G;
        }
        else
        {
        // critical path: partition 0 against the newest spectrum (direct loads: no shared memory, so its CTAs fit
        // beside the tail's on every SM)
        Range rh = c->r_head;
        rh.slot = c->g.slot; rh.kind = 1;
        if ((r = launch_cmac<T>(c, rh, c->d_S.p, 0, st))) return r;
        if (pe) HB_CUDA(cudaEventRecord(pe[2], st));
        sets.n = 1;
        sets.s[0].S = c->d_S.p; sets.s[0].U = rh.U; sets.s[0].upt = rh.upt; sets.s[0].G = rh.G;
        // join: the tail of THIS hop was launched during the previous one
        if (had_tail)
        {
            HB_CUDA(cudaStreamWaitEvent(st, c->ev_tail[old_par], 0));
            sets.n = 2;
            sets.s[1].S = c->d_St[old_par].p; sets.s[1].U = c->r_tail.U; sets.s[1].upt = c->r_tail.upt; sets.s[1].G = c->r_tail.G;
        }
        if (pe) HB_CUDA(cudaEventRecord(pe[3], st));
        }
    }
    if ((r = launch_inv<T>(c, sets, io, st, peer))) return r;
    if (pe) HB_CUDA(cudaEventRecord(pe[4], st));
    return HB_OK;
}

// core of process: device rows in, device rows out, everything enqueued on st
template <class T>
int process_core(hb_conv *c, const T *d_in, size_t in_ld, T *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    if (!c->P) return HB_ERR_NO_IR;
    if (!n) return HB_OK;
    int rc;
    if ((rc = ensure_ready<T>(c, n, st))) return rc;
    const size_t B = c->g.B;
    const size_t rows_in = size_t(c->groups) * c->ins, rows_out = size_t(c->groups) * c->outs;
    const size_t rw = c->rw;
    const size_t nh = (rw + n) / B;
    auto hop = [&](const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld, const InvIO<T> &io) -> int
    {
        return launch_hop<T>(c, st, prev, prev_ld, newest, new_ld, save, save_ld, io, PeerOut());
    };

    // caller rows that overlap (in-place processing) must go through the staging copies
    const char *ib = (const char *) d_in, *ie = (const char *) (d_in + (rows_in - 1) * in_ld + n);
    const char *ob = (const char *) d_out, *oe = (const char *) (d_out + (rows_out - 1) * out_ld + n);
    const bool disjoint = !d_out || ie <= ob || oe <= ib;
    // d_out == nullptr: the caller (deferred host path) fetches the finished block itself
    if (rw == 0 && nh * B == n && disjoint && (d_out || nh <= 1) && c->xin_ld >= B && c->yout_ld >= B)
    {
        // Hop-aligned call (the usual audio-callback case): no staging copies.  The forward kernel reads the
        // caller's rows directly and saves the last block as the next call's "previous hop"; the inverse
        // kernel of the first hop hands the block finished by the previous call to the caller (the output is
        // the convolution delayed by exactly B) and the last hop's block stays behind for the next call.
        const int cur = c->cur, nxt = cur ^ 1;
        const T *x_keep = (const T *) c->d_xin[cur].p + c->x_tail;
        const T *y_keep = (const T *) c->d_yout[cur].p + c->y_tail;
        // where hop h of this call reads its frame and leaves its block
        auto inv_io = [&](size_t h) -> InvIO<T>
        {
            const bool first = h == 0, last = h + 1 == nh;
            InvIO<T> io;
            io.yout = last ? (T *) c->d_yout[nxt].p : d_out; io.ld = last ? c->yout_ld : out_ld; io.off = last ? 0 : (h + 1) * B;
            io.add_result = last ? 0 : accumulate;
            io.carry_src = first ? y_keep : nullptr; io.carry_src_ld = c->yout_ld;
            io.carry_dst = (first && d_out) ? d_out : nullptr; io.carry_dst_ld = out_ld; io.add_carry = accumulate;
            return io;
        };
        size_t h = 0;
        while (h < nh)
        {
            const size_t left = nh - h;
            int nb = 1;
            const bool quiet = c->reveals.empty();                    // restarted pairs come back one partition per hop: no batches meanwhile
            while (quiet && nb * 2 <= c->mh_max && size_t(nb) * 2 <= left) nb *= 2;
            const bool small_batch = quiet && nb == 1 && c->hb_max > 1 && left >= 2;
            if (small_batch) nb = (int) std::min<size_t>(left, c->hb_max);
            if (nb == 1)
            {
                const bool first = h == 0, last = h + 1 == nh;
                rc = hop(first ? x_keep : d_in + (h - 1) * B, first ? c->xin_ld : in_ld, d_in + h * B, in_ld,
                         last ? (T *) c->d_xin[nxt].p : nullptr, c->xin_ld, inv_io(h));
                if (rc) return rc;
                h++;
                continue;
            }
            // ---- multi-hop reuse: nb hops over ONE pass of the IR spectra (k_cmac_tma_mh) ----
            c->chain_ok = false;
            // a tail launched ahead for the first of these hops is not used: the batch covers all partitions itself
            if (c->split && c->tail_valid) HB_CUDA(cudaStreamWaitEvent(st, c->ev_tail[c->tail_par], 0));
            c->tail_valid = false;
            // forward transforms of the nb hops in one launch: hop j = block h + j of the caller's rows into slot slot0 - j
            const bool first = h == 0, last_in_batch = h + nb == nh;
            c->g.slot = c->g.slot ? c->g.slot - 1 : c->g.R - 1;
            const uint32_t slot0 = c->g.slot;
            c->g.hop += nb;
            c->g.trace = (unsigned long long *) c->d_trace.p;
            if ((rc = launch_fwd<T>(c, first ? x_keep : d_in + (h - 1) * B, first ? c->xin_ld : in_ld, d_in + h * B, in_ld,
                                    last_in_batch ? (T *) c->d_xin[nxt].p : nullptr, c->xin_ld, st, (uint32_t) nb))) return rc;
            Range rf = c->r_full;
            rf.slot = slot0; rf.kind = 2;
            const uint32_t split = small_batch ? 1u : mh_split(c, nb);
            const uint64_t set_stride = uint64_t(rf.G + c->g.tiles * split) * (c->g.Q / split);
            if ((rc = c->d_Smh.ensure(size_t(std::max<uint32_t>(MH_MAX, c->hb_max)) * uint64_t(rf.G + c->g.tiles) * c->g.Q * 16))) return rc;
            if (small_batch)
            {
                // launch-latency-bound engine: nothing to reuse (its spectra sit in L2), the hops just share their launches
                if ((rc = launch_cmac<T>(c, rf, c->d_Smh.p, 0, st, (uint32_t) nb, set_stride))) return rc;
            }
            else if ((rc = launch_cmac_mh<T>(c, rf, c->d_Smh.p, nb, set_stride, st))) return rc;
            // inverse transforms of the nb hops in one launch: hop j sums set j, runs its Nyquist products from slot0 - j and
            // leaves its block B samples further on; the last hop of the call stays behind in the staging row
            {
                SegSets sets;
                memset(&sets, 0, sizeof(sets));
                sets.n = 1;
                sets.s[0].S = c->d_Smh.p; sets.s[0].U = rf.U * split; sets.s[0].upt = rf.upt; sets.s[0].G = rf.G; sets.s[0].split = split;
                InvIO<T> io;
                io.yout = d_out; io.ld = out_ld; io.off = (h + 1) * B; io.add_result = accumulate;
                io.carry_src = first ? y_keep : nullptr; io.carry_src_ld = c->yout_ld;
                io.carry_dst = (first && d_out) ? d_out : nullptr; io.carry_dst_ld = out_ld; io.add_carry = accumulate;
                InvBatch ib = no_batch();
                ib.set_stride = set_stride;
                if (last_in_batch) { ib.last_j = nb - 1; ib.last_yout = c->d_yout[nxt].p; ib.last_ld = c->yout_ld; }
                if ((rc = launch_inv<T>(c, sets, io, st, PeerOut(), (uint32_t) nb, ib))) return rc;
            }
            // the newest spectrum now sits nb - 1 slots below slot0
            c->g.slot = slot0 >= uint32_t(nb - 1) ? slot0 - (nb - 1) : slot0 + c->g.R - (nb - 1);
            // nothing is launched ahead for the hop after a batch (the next call is most likely another batch, which
            // would discard it): a following single hop of the overlapped schedule runs all its partitions itself
            if (c->split) c->tail_missing = true;
            h += nb;
        }
        c->cur = nxt;
        c->x_tail = c->y_tail = 0;
        return HB_OK;
    }

    if ((rc = ensure_staging<T>(c, B + rw + n, (nh + 1) * B, true, st))) return rc;
    c->chain_ok = false;                                              // row copies sit between the hops of this path

    // bring the retained heads to offset 0 of the other row set when they sit further in
    if (c->x_tail || c->y_tail)
    {
        const int nxt = c->cur ^ 1;
        if ((rc = launch_rows<T>((T *) c->d_xin[nxt].p, c->xin_ld, (const T *) c->d_xin[c->cur].p + c->x_tail, c->xin_ld, B + rw, rows_in, 0, st))) return rc;
        if ((rc = launch_rows<T>((T *) c->d_yout[nxt].p, c->yout_ld, (const T *) c->d_yout[c->cur].p + c->y_tail, c->yout_ld, B, rows_out, 0, st))) return rc;
        c->cur = nxt;
        c->x_tail = c->y_tail = 0;
    }
    T *xin = (T *) c->d_xin[c->cur].p;
    T *yout = (T *) c->d_yout[c->cur].p;
    if ((rc = launch_rows<T>(xin + B + rw, c->xin_ld, d_in, in_ld, n, rows_in, 0, st))) return rc;

    for (size_t h = 0; h < nh; h++)
    {
        InvIO<T> io = {yout, c->yout_ld, (h + 1) * B, 0, nullptr, 0, nullptr, 0, 0};
        if ((rc = hop(xin + h * B, c->xin_ld, xin + (h + 1) * B, c->xin_ld, nullptr, 0, io))) return rc;
    }
    if (d_out && (rc = launch_rows<T>(d_out, out_ld, yout + rw, c->yout_ld, n, rows_out, accumulate, st))) return rc;
    c->chain_ok = false;

    c->rw = (rw + n) - nh * B;
    c->x_tail = nh * B;
    c->y_tail = nh * B;
    return HB_OK;
}

int core_dispatch(hb_conv *c, const void *d_in, size_t in_ld, void *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    return c->dtype == HB_F64 ? process_core<double>(c, (const double *) d_in, in_ld, (double *) d_out, out_ld, n, accumulate, st)
                              : process_core<float>(c, (const float *) d_in, in_ld, (float *) d_out, out_ld, n, accumulate, st);
}

// how long k_gather waits for a peer's blocks before it raises the late flag (HB_PEER_TIMEOUT_MS, default 30 s: a peer may
// still be loading impulse responses)
unsigned long long peer_timeout_ns()
{
    static const char *env = getenv("HB_PEER_TIMEOUT_MS");
    const double ms = env && atof(env) > 0 ? atof(env) : 30000.0;
    return (unsigned long long) (ms * 1e6);
}

// A call of a rank of the fused multi-GPU exchange: as process_core, but the inverse kernel of every hop delivers each
// partial block into its owner's inbox and k_gather sums what arrived here.  d_out holds this rank's outs/world output
// rows (nullptr: the caller fetches the finished block itself -- deferred host path, at most one hop).  Any numSamples:
// hop-aligned calls run straight from / into the caller's rows, the others through the staging rows.
template <class T>
int process_shard(hb_conv *c, const T *d_in, size_t in_ld, T *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    // a rank without an impulse response still takes part (k_shard_silence): its peers wait for its blocks
    const bool silent = c->P == 0;
    if (!n) return HB_OK;
    int rc;
    if ((rc = ensure_ready<T>(c, n, st))) return rc;
    const size_t B = c->g.B;
    const uint32_t world = c->shard_world, o_loc = c->outs / world;
    const size_t rows_in = c->ins, rows_out = o_loc;
    const size_t rw = c->rw;
    const size_t nh = (rw + n) / B;

    // the exchange of nb consecutive hops starting with the next sequence number: where every partial block goes ...
    auto make_peer = [&](uint32_t nb, GatherBatch &gb) -> PeerOut
    {
        PeerOut peer;
        for (uint32_t r = 0; r < world; r++)
        {
            peer.data[r] = c->peer_base[r];
            peer.count[r] = (uint32_t *) ((char *) c->peer_base[r] + c->inbox_data_bytes);
        }
        peer.world = world; peer.rank = c->shard_rank; peer.outs_local = o_loc;
        peer.parity = (c->hop_seq + 1) % HB_INBOX_DEPTH;
        peer.slot = c->inbox_slot;
        memset(&gb, 0, sizeof(gb));
        gb.last_j = -1;
        for (uint32_t j = 0; j < nb; j++)
        {
            const uint32_t par = (++c->hop_seq) % HB_INBOX_DEPTH;
            gb.e[j] = (++c->parity_uses[par]) * o_loc;
        }
        return peer;
    };
    // ... and the owner-side sum of the blocks that arrived here (nb hops in one launch)
    auto gather = [&](const PeerOut &peer, const GatherBatch &gb, uint32_t nb, T *yout, size_t ld, size_t off, int add_result,
                      const T *carry_src, T *carry_dst, size_t carry_dst_ld) -> int
    {
        k_gather<T><<<dim3(o_loc, nb), 256, 0, st>>>((const T *) c->d_inbox, (const uint32_t *) ((const char *) c->d_inbox + c->inbox_data_bytes), world, o_loc,
                                                     peer.parity, peer.slot, gb, (uint32_t) B, yout, ld, off, add_result,
                                                     carry_src, c->yout_ld, carry_dst, carry_dst_ld, accumulate,
                                                     (unsigned long long *) c->d_trace.p, c->g.hop, peer_timeout_ns(),
                                                     (uint32_t *) ((char *) c->d_inbox + c->inbox_data_bytes + INBOX_LATE_OFF));
        HB_LAUNCH_CHECK();
        return HB_OK;
    };
    // one hop: this rank's partial blocks to their owners, then the owner-side sum of the blocks that arrived here into
    // row set `yout` at offset `off`; carry_dst (optional) first receives the block at carry_src (the output-ring read)
    auto shard_hop = [&](const T *prev, size_t prev_ld, const T *newest, size_t new_ld, T *save, size_t save_ld,
                         T *yout, size_t ld, size_t off, int add_result, const T *carry_src, T *carry_dst, size_t carry_dst_ld) -> int
    {
        GatherBatch gb;
        const PeerOut peer = make_peer(1, gb);
        InvIO<T> io = {nullptr, 0, 0, 0, nullptr, 0, nullptr, 0, 0};
        int r;
        if (silent)
        {
            k_shard_silence<T><<<c->outs, 256, 0, st>>>(peer, (uint32_t) B);
            HB_LAUNCH_CHECK();
        }
        else if ((r = launch_hop<T>(c, st, prev, prev_ld, newest, new_ld, save, save_ld, io, peer))) return r;
        return gather(peer, gb, 1, yout, ld, off, add_result, carry_src, carry_dst, carry_dst_ld);
    };

    const char *ib = (const char *) d_in, *ie = (const char *) (d_in + (rows_in - 1) * in_ld + n);
    const char *ob = (const char *) d_out, *oe = (const char *) (d_out + (rows_out - 1) * out_ld + n);
    const bool disjoint = !d_out || ie <= ob || oe <= ib;
    if (rw == 0 && nh * B == n && disjoint && (d_out || nh <= 1) && c->xin_ld >= B && c->yout_ld >= B)
    {
        const int cur = c->cur, nxt = cur ^ 1;
        const T *x_keep = (const T *) c->d_xin[cur].p + c->x_tail;
        const T *y_keep = (const T *) c->d_yout[cur].p + c->y_tail;
        size_t h = 0;
        while (h < nh)
        {
            const bool first = h == 0;
            int nb = 1;
            if (!silent && c->mh_ok && c->reveals.empty()) while (nb * 2 <= c->mh_max && size_t(nb) * 2 <= nh - h) nb *= 2;
            if (nb == 1)
            {
                const bool last = h + 1 == nh;
                if ((rc = shard_hop(first ? x_keep : d_in + (h - 1) * B, first ? c->xin_ld : in_ld, d_in + h * B, in_ld,
                                    last ? (T *) c->d_xin[nxt].p : nullptr, c->xin_ld,
                                    last ? (T *) c->d_yout[nxt].p : d_out, last ? c->yout_ld : out_ld, last ? 0 : (h + 1) * B, last ? 0 : accumulate,
                                    first ? y_keep : nullptr, (first && d_out) ? d_out : nullptr, out_ld))) return rc;
                h++;
                continue;
            }
            // ---- multi-hop reuse on the sharded engine: nb hops over ONE pass of this rank's IR spectra, their partial blocks
            // delivered by one inverse launch (hop j into inbox slot seq + j), summed by one owner-side launch ----
            if (c->split && c->tail_valid) HB_CUDA(cudaStreamWaitEvent(st, c->ev_tail[c->tail_par], 0));
            c->tail_valid = false;
            const bool last_in_batch = h + nb == nh;
            c->g.slot = c->g.slot ? c->g.slot - 1 : c->g.R - 1;
            const uint32_t slot0 = c->g.slot;
            c->g.hop += nb;
            c->g.trace = (unsigned long long *) c->d_trace.p;
            if ((rc = launch_fwd<T>(c, first ? x_keep : d_in + (h - 1) * B, first ? c->xin_ld : in_ld, d_in + h * B, in_ld,
                                    last_in_batch ? (T *) c->d_xin[nxt].p : nullptr, c->xin_ld, st, (uint32_t) nb))) return rc;
            Range rf = c->r_full;
            rf.slot = slot0; rf.kind = 2;
            const uint32_t split = mh_split(c, nb);
            const uint64_t set_stride = uint64_t(rf.G + c->g.tiles * split) * (c->g.Q / split);
            if ((rc = c->d_Smh.ensure(size_t(MH_MAX) * uint64_t(rf.G + c->g.tiles) * c->g.Q * 16))) return rc;
            if ((rc = launch_cmac_mh<T>(c, rf, c->d_Smh.p, nb, set_stride, st))) return rc;
            GatherBatch gb;
            const PeerOut peer = make_peer((uint32_t) nb, gb);
            {
                SegSets sets;
                memset(&sets, 0, sizeof(sets));
                sets.n = 1;
                sets.s[0].S = c->d_Smh.p; sets.s[0].U = rf.U * split; sets.s[0].upt = rf.upt; sets.s[0].G = rf.G; sets.s[0].split = split;
                InvIO<T> io = {nullptr, 0, 0, 0, nullptr, 0, nullptr, 0, 0};
                InvBatch ib = no_batch();
                ib.set_stride = set_stride;
                if ((rc = launch_inv<T>(c, sets, io, st, peer, (uint32_t) nb, ib))) return rc;
            }
            if (last_in_batch) { gb.last_j = nb - 1; gb.last_yout = c->d_yout[nxt].p; gb.last_ld = c->yout_ld; }
            if ((rc = gather(peer, gb, (uint32_t) nb, d_out, out_ld, (h + 1) * B, accumulate, first ? y_keep : nullptr, (first && d_out) ? d_out : nullptr, out_ld))) return rc;
            c->g.slot = slot0 >= uint32_t(nb - 1) ? slot0 - (nb - 1) : slot0 + c->g.R - (nb - 1);
            if (c->split) c->tail_missing = true;
            h += nb;
        }
        c->cur = nxt;
        c->x_tail = c->y_tail = 0;
        return HB_OK;
    }

    if ((rc = ensure_staging<T>(c, B + rw + n, (nh + 1) * B, true, st))) return rc;
    if (c->x_tail || c->y_tail)
    {
        const int nxt = c->cur ^ 1;
        if ((rc = launch_rows<T>((T *) c->d_xin[nxt].p, c->xin_ld, (const T *) c->d_xin[c->cur].p + c->x_tail, c->xin_ld, B + rw, rows_in, 0, st))) return rc;
        if ((rc = launch_rows<T>((T *) c->d_yout[nxt].p, c->yout_ld, (const T *) c->d_yout[c->cur].p + c->y_tail, c->yout_ld, B, rows_out, 0, st))) return rc;
        c->cur = nxt;
        c->x_tail = c->y_tail = 0;
    }
    T *xin = (T *) c->d_xin[c->cur].p;
    T *yout = (T *) c->d_yout[c->cur].p;
    if ((rc = launch_rows<T>(xin + B + rw, c->xin_ld, d_in, in_ld, n, rows_in, 0, st))) return rc;
    for (size_t h = 0; h < nh; h++)
        if ((rc = shard_hop(xin + h * B, c->xin_ld, xin + (h + 1) * B, c->xin_ld, nullptr, 0, yout, c->yout_ld, (h + 1) * B, 0, nullptr, nullptr, 0))) return rc;
    if (d_out && (rc = launch_rows<T>(d_out, out_ld, yout + rw, c->yout_ld, n, rows_out, accumulate, st))) return rc;
    c->rw = (rw + n) - nh * B;
    c->x_tail = nh * B;
    c->y_tail = nh * B;
    return HB_OK;
}

int shard_dispatch(hb_conv *c, const void *d_in, size_t in_ld, void *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    return c->dtype == HB_F64 ? process_shard<double>(c, (const double *) d_in, in_ld, (double *) d_out, out_ld, n, accumulate, st)
                              : process_shard<float>(c, (const float *) d_in, in_ld, (float *) d_out, out_ld, n, accumulate, st);
}

// rows a host-pointer call brings and takes: a rank of the fused exchange returns only the outputs it owns
inline size_t host_rows_in(const hb_conv *c) { return size_t(c->groups) * c->ins; }
inline size_t host_rows_out(const hb_conv *c) { return c->peers_attached ? c->outs / c->shard_world : size_t(c->groups) * c->outs; }
// the engine behind a host-pointer call: the local matrix, or this rank's share of the fused exchange
inline int host_dispatch(hb_conv *c, const void *d_in, size_t in_ld, void *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    return c->peers_attached ? shard_dispatch(c, d_in, in_ld, d_out, out_ld, n, accumulate, st)
                             : core_dispatch(c, d_in, in_ld, d_out, out_ld, n, accumulate, st);
}

// wait for the copy streams of the deferred host path (before anything else touches the staging rows)
int drain_deferred(hb_conv *c)
{
    if (c->s_h2d) HB_CUDA(cudaStreamSynchronize(c->s_h2d));
    if (c->s_d2h) HB_CUDA(cudaStreamSynchronize(c->s_d2h));
    for (int k = 0; k < 2; k++) c->blk_pending[k] = c->inq_pending[k] = c->done_pending[k] = false;
    return HB_OK;
}

template <class T>
int set_ir_core(hb_conv *c, uint32_t grp, uint32_t in, uint32_t o, const T *d_ir, uintptr_t length, cudaStream_t st)
{
    // length clipping of PartitionedConvolve::set (cpp:192-199); d_ir already points at tap `offset`
    int error = ERR_NONE;
    if (c->length && c->length < length) length = c->length;
    if (length > c->max_length) { length = c->max_length; error = ERR_MEM_ALLOC_TOO_SMALL; }
    plan_geometry(c);
    const uint32_t B = c->g.B;
    const uint32_t np = (uint32_t) ((length + B - 1) / B);
    const size_t pair = (size_t(grp) * c->outs + o) * c->ins + in;
    const uint32_t nwrite = std::max(np, c->nparts[pair]);
    drop_reveal(c, grp, in, o);
    int rc = launch_ir<T>(c, d_ir, length, grp, in, o, nwrite, st);
    if (rc) return rc;
    c->nparts[pair] = np;
    c->P = *std::max_element(c->nparts.begin(), c->nparts.end());
    c->need_reset = true;
    return error;
}

// Can pair restarts be done in place?  The stream must be running (otherwise a plain set / reset is the same thing), the pair must fit
// the delay line as it is (a longer ring is a new geometry) and the transforms must be single-CTA ones (k_ir's side output).
template <class T>
bool live_possible(hb_conv *c, uint32_t np)
{
    return !c->need_reset && c->P > 0 && np <= c->P && !is_big<T>(c);
}

// hide the pair and register its spectra for the partition-per-hop return; `from_ir`: transform d_ir (nullptr / 0 taps = clear the pair),
// else gather the spectra it has now (reset of the pair)
template <class T>
int restart_pair(hb_conv *c, uint32_t grp, uint32_t in, uint32_t o, bool from_ir, const T *d_ir, uintptr_t length, uint32_t np)
{
    // the hops in flight read the spectra that are about to change
    HB_CUDA(cudaDeviceSynchronize());
    const size_t pair = (size_t(grp) * c->outs + o) * c->ins + in;
    const uint32_t old_np = c->nparts[pair];
    drop_reveal(c, grp, in, o);
    if (c->split) { c->tail_valid = false; c->tail_missing = true; }      // the tail computed ahead holds the old pair
    hb_conv::Reveal r;
    r.grp = grp; r.in = in; r.out = o; r.np = np; r.shown = 0; r.age = 0; r.side = nullptr;
    int rc = HB_OK;
    if (np)
    {
        HB_CUDA(cudaMalloc(&r.side, size_t(np) * c->g.B * sizeof(Cx<T>) + size_t(np) * sizeof(T)));
        if (from_ir) rc = launch_ir<T>(c, d_ir, length, grp, in, o, np, c->stream, r.side);
        else rc = pair_copy<T>(c, r, 0, np, 1, c->stream);
    }
    if (rc == HB_OK) rc = pair_copy<T>(c, r, 0, std::max(old_np, np), 2, c->stream);
    if (rc) { cudaFree(r.side); return rc; }
    HB_CUDA(cudaStreamSynchronize(c->stream));
    c->nparts[pair] = np;
    if (np) c->reveals.push_back(r);
    return HB_OK;
}

template <class T>
int set_ir_live_core(hb_conv *c, uint32_t grp, uint32_t in, uint32_t o, const T *d_ir, uintptr_t length)
{
    int error = ERR_NONE;
    if (c->length && c->length < length) length = c->length;
    if (length > c->max_length) { length = c->max_length; error = ERR_MEM_ALLOC_TOO_SMALL; }
    const uint32_t B = c->g.B;
    const uint32_t np = (uint32_t) ((length + B - 1) / B);
    if (!live_possible<T>(c, np))
    {
        int rc = set_ir_core<T>(c, grp, in, o, d_ir, length, c->stream);
        return rc < 0 ? rc : (rc ? rc : error);
    }
    int rc = restart_pair<T>(c, grp, in, o, true, d_ir, length, np);
    return rc ? rc : error;
}

int check_handle(hb_conv *c)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    return use_device(c->device);
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int hb_conv_create(hb_conv **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs,
                              uintptr_t max_fft_size, uintptr_t max_length, uintptr_t offset, uintptr_t length, int device)
{
    if (!out || (dtype != HB_F32 && dtype != HB_F64) || !groups || !ins || !outs) { set_error("hb_conv_create: bad argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = use_device(device);
    if (rc) return rc;

    // setMaxFFTSize (PartitionedConvolve.cpp:26-50)
    int ctor_error = ERR_NONE;
    uintptr_t l2 = ceil_log2(max_fft_size);
    if (l2 > MAX_FFT_LOG2) { ctor_error = ERR_FFT_SIZE_MAX_TOO_LARGE; l2 = MAX_FFT_LOG2; }
    if (l2 && l2 < MIN_FFT_LOG2) { ctor_error = ERR_FFT_SIZE_MAX_TOO_SMALL; l2 = MIN_FFT_LOG2; }
    if (max_fft_size != (uintptr_t(1) << l2)) ctor_error = ERR_FFT_SIZE_MAX_NON_POWER_OF_TWO;
    if (l2 < MIN_FFT_LOG2) l2 = MIN_FFT_LOG2;

    hb_conv *c = new hb_conv;
    c->dtype = dtype; c->device = device; c->groups = groups; c->ins = ins; c->outs = outs;
    c->max_fft_log2 = l2;
    if (!fft_supported(c, l2))
    {
        set_error("FFT size 2^%d is beyond what this build implements (2^%d)", (int) l2, BIG_MAX_LOG2 + 1);
        delete c;
        return HB_ERR_UNSUPPORTED;
    }
    c->fft_log2 = l2;
    c->offset = offset;
    const uintptr_t maxB = (uintptr_t(1) << l2) >> 1;
    uintptr_t ml = max_length ? max_length : maxB;
    if (ml % maxB) ml = (ml / maxB + 1) * maxB;                 // cpp:77-82
    c->max_length = ml;
    c->length = std::min(length, ml);                           // setLength (cpp:156-161)
    c->nparts.assign(c->pairs(), 0u);

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete c; return HB_ERR_CUDA; }
    static const char *env_ts = getenv("HB_TAIL_STREAMS");           // experiments: default number of tail streams
    if (env_ts && (atoi(env_ts) == 1 || atoi(env_ts) == 2)) c->tail_streams_req = atoi(env_ts);
    c->tw_log2 = (int) l2;
    rc = make_twiddles(dtype, c->tw_log2, &c->d_tw);
    if (rc == HB_OK) rc = alloc_capacity(c);
    if (rc != HB_OK && rc != ERR_NONE)
    {
        free_device(c);
        cudaStreamDestroy(c->stream);
        delete c;
        return rc;
    }
    plan_geometry(c);
    *out = c;
    return ctor_error;
}

extern "C" void hb_conv_destroy(hb_conv *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_device(c);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int hb_conv_set_fft_size(hb_conv *c, uintptr_t fft_size)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    return apply_fft_size(c, fft_size);
}

extern "C" int hb_conv_set_length(hb_conv *c, uintptr_t length)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->length = std::min(length, c->max_length);
    return length > c->max_length ? ERR_PARTITION_LENGTH_TOO_LARGE : ERR_NONE;
}

extern "C" int hb_conv_set_offset(hb_conv *c, uintptr_t offset)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->offset = offset;
    return ERR_NONE;
}

extern "C" int hb_conv_set_reset_offset(hb_conv *c, intptr_t offset)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->reset_offset = offset;
    return ERR_NONE;
}

namespace
{
int set_ir_host(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length, bool live);
}

extern "C" int hb_conv_set_ir(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length)
{
    return set_ir_host(c, group, in, out, ir, ir_dtype, length, false);
}

extern "C" int hb_conv_set_ir_live(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length)
{
    return set_ir_host(c, group, in, out, ir, ir_dtype, length, true);
}

extern "C" int hb_conv_reset_pair(hb_conv *c, uint32_t group, uint32_t in, uint32_t out)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (group >= c->groups || in >= c->ins || out >= c->outs) { set_error("hb_conv_reset_pair: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    const size_t pair = (size_t(group) * c->outs + out) * c->ins + in;
    const uint32_t np = c->nparts[pair];
    const bool live = c->dtype == HB_F64 ? live_possible<double>(c, np) : live_possible<float>(c, np);
    if (!live) { c->need_reset = true; return ERR_NONE; }              // nothing running yet, or not possible in place: the whole stream restarts
    if (!np) return ERR_NONE;
    // a pair that is still coming back keeps the spectra it has not shown yet: put them in place before they are gathered again
    for (hb_conv::Reveal &r : c->reveals)
        if (r.grp == group && r.in == in && r.out == out)
        {
            rc = c->dtype == HB_F64 ? pair_copy<double>(c, r, r.shown, r.np - r.shown, 0, c->stream) : pair_copy<float>(c, r, r.shown, r.np - r.shown, 0, c->stream);
            if (rc) return rc;
            HB_CUDA(cudaStreamSynchronize(c->stream));
        }
    return c->dtype == HB_F64 ? restart_pair<double>(c, group, in, out, false, nullptr, 0, np) : restart_pair<float>(c, group, in, out, false, nullptr, 0, np);
}

namespace
{
int set_ir_host(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length, bool live)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (group >= c->groups || in >= c->ins || out >= c->outs || (ir_dtype != HB_F32 && ir_dtype != HB_F64)) { set_error("hb_conv_set_ir: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    // cpp:192: nothing to load when the IR ends before the offset
    uintptr_t len = (!ir || length <= c->offset) ? 0 : length - c->offset;
    uintptr_t take = len;
    if (c->length && c->length < take) take = c->length;
    if (take > c->max_length) take = c->max_length;
    const size_t es = c->esize();
    if (take)
    {
        if ((rc = c->h_ir.ensure(take * es)) || (rc = c->d_ir.ensure(take * es))) return rc;
        // the previous upload from the pinned buffer must have landed before it is overwritten
        HB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->dtype == HB_F64)
        {
            double *dst = (double *) c->h_ir.p;
            if (ir_dtype == HB_F64) memcpy(dst, (const double *) ir + c->offset, take * sizeof(double));
            else for (size_t k = 0; k < take; k++) dst[k] = (double) ((const float *) ir)[c->offset + k];
        }
        else
        {
            float *dst = (float *) c->h_ir.p;
            if (ir_dtype == HB_F32) memcpy(dst, (const float *) ir + c->offset, take * sizeof(float));
            else for (size_t k = 0; k < take; k++) dst[k] = (float) ((const double *) ir)[c->offset + k];   // Convolver.cpp:126-134
        }
        HB_CUDA(cudaMemcpyAsync(c->d_ir.p, c->h_ir.p, take * es, cudaMemcpyHostToDevice, c->stream));
    }
    if (live)
        rc = c->dtype == HB_F64 ? set_ir_live_core<double>(c, group, in, out, (const double *) c->d_ir.p, len)
                                : set_ir_live_core<float>(c, group, in, out, (const float *) c->d_ir.p, len);
    else
        rc = c->dtype == HB_F64 ? set_ir_core<double>(c, group, in, out, (const double *) c->d_ir.p, len, c->stream)
                                : set_ir_core<float>(c, group, in, out, (const float *) c->d_ir.p, len, c->stream);
    if (rc < 0) return rc;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}
} // namespace

extern "C" int hb_conv_set_ir_dev(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *d_ir, uintptr_t length)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (group >= c->groups || in >= c->ins || out >= c->outs) { set_error("hb_conv_set_ir_dev: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    // d_ir may have been produced on any stream of the caller (the transforms run on the engine's own non-blocking
    // stream, which is not ordered after any of them): wait for everything enqueued on the device so far
    HB_CUDA(cudaDeviceSynchronize());
    uintptr_t len = (!d_ir || length <= c->offset) ? 0 : length - c->offset;
    const char *p = (const char *) d_ir + (len ? c->offset * c->esize() : 0);
    rc = c->dtype == HB_F64 ? set_ir_core<double>(c, group, in, out, (const double *) p, len, c->stream)
                            : set_ir_core<float>(c, group, in, out, (const float *) p, len, c->stream);
    if (rc < 0) return rc;
    // as hb_conv_set_ir: d_ir may be reused and any stream may process once this returns
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

extern "C" int hb_conv_resize(hb_conv *c, uintptr_t max_length)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->s_tail) HB_CUDA(cudaStreamSynchronize(c->s_tail));       // a tail launched ahead reads the spectra freed below
    if (c->s_tail_b) HB_CUDA(cudaStreamSynchronize(c->s_tail_b));
    const uintptr_t maxB = (uintptr_t(1) << c->max_fft_log2) >> 1;
    uintptr_t ml = max_length ? max_length : maxB;
    if (ml % maxB) ml = (ml / maxB + 1) * maxB;
    if (ml == c->max_length) return ERR_NONE;

    // keep what is loaded: same tiling, only the partition stride (Pcap) of the layout changes
    plan_geometry(c);
    if ((rc = c->dtype == HB_F64 ? finish_reveals<double>(c, c->stream) : finish_reveals<float>(c, c->stream))) return rc;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    const Geom old = c->g;
    void *oldH = c->d_H, *oldHn = c->d_Hnyq;
    const std::vector<uint32_t> old_parts = c->nparts;
    c->d_H = c->d_Hnyq = nullptr;
    c->max_length = ml;
    if (c->length > ml) c->length = ml;
    rc = alloc_capacity(c);
    if (rc)
    {
        cudaFree(oldH); cudaFree(oldHn);
        return rc;
    }
    plan_geometry(c);
    const Geom &now = c->g;
    const uint32_t keep = std::min(old.Pcap, now.Pcap);
    if (keep && oldH)
    {
        const size_t unit = size_t(old.Q) * 16;
        HB_CUDA(cudaMemcpy2DAsync(c->d_H, size_t(now.Pcap) * unit, oldH, size_t(old.Pcap) * unit, size_t(keep) * unit,
                                  size_t(old.tiles) * old.ins, cudaMemcpyDeviceToDevice, c->stream));
        HB_CUDA(cudaMemcpy2DAsync(c->d_Hnyq, size_t(now.Pcap) * c->esize(), oldHn, size_t(old.Pcap) * c->esize(), size_t(keep) * c->esize(),
                                  c->pairs(), cudaMemcpyDeviceToDevice, c->stream));
        for (size_t k = 0; k < c->nparts.size(); k++) c->nparts[k] = std::min(old_parts[k], keep);
        c->P = *std::max_element(c->nparts.begin(), c->nparts.end());
    }
    HB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(oldH); cudaFree(oldHn);
    c->need_reset = true;
    return ERR_NONE;
}

extern "C" int hb_conv_reset(hb_conv *c)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->need_reset = true;
    return ERR_NONE;
}

extern "C" int hb_conv_dtype(const hb_conv *c) { return c ? c->dtype : HB_F32; }
extern "C" uintptr_t hb_conv_partitions(const hb_conv *c) { return c ? c->P : 0; }
extern "C" uintptr_t hb_conv_max_length(const hb_conv *c) { return c ? c->max_length : 0; }
extern "C" uintptr_t hb_conv_fft_size(const hb_conv *c) { return c ? (uintptr_t(1) << c->fft_log2) : 0; }

extern "C" int hb_conv_process_dev(hb_conv *c, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld,
                                   uintptr_t num_samples, int accumulate, void *stream)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if ((!d_in || !d_out) && num_samples) { set_error("hb_conv_process_dev: null buffer"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
    if (c->blk_valid)
    {
        c->chain_ok = false;
        if ((rc = drain_deferred(c))) return rc;                    // host-pointer calls were in flight on this handle
    }
    c->blk_valid = false;
    return core_dispatch(c, d_in, in_ld, d_out, out_ld, num_samples, accumulate, st);
}

namespace
{
// Row copies between the caller's arrays and the pinned staging buffers are the host cost of a large call (config 5: 1 MiB each
// way per 90 us hop -- a single thread's memcpy alone takes longer than the hop).  Calls that move 256 KiB or more share their rows
// out to a few helper threads (started on first use, spinning briefly between calls of a streaming caller, asleep otherwise).
class RowCopyPool
{
public:
    static RowCopyPool &get() { static RowCopyPool p; return p; }
    // fn(r) for r in [0, rows), rows dealt to the helpers and the caller
    void run(size_t rows, size_t bytes, const std::function<void(size_t)> &fn)
    {
        if (bytes < (size_t(256) << 10) || rows < 2 || !helpers_) { for (size_t r = 0; r < rows; r++) fn(r); return; }
        std::lock_guard<std::mutex> one(call_);                      // one call at a time owns the helpers
        fn_ = &fn;
        rows_ = rows;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(helpers_, std::memory_order_release);
        {
            std::lock_guard<std::mutex> l(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        while (pending_.load(std::memory_order_acquire) != 0) cpu_pause();
    }
    ~RowCopyPool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_.store(true, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }

private:
    RowCopyPool()
    {
        const unsigned hw = std::thread::hardware_concurrency();
        helpers_ = hw >= 8 ? 3 : (hw >= 4 ? 1 : 0);
        for (unsigned k = 0; k < helpers_; k++) threads_.emplace_back([this]() { loop(); });
    }
    static void cpu_pause()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void work()
    {
        for (size_t r; (r = next_.fetch_add(1, std::memory_order_relaxed)) < rows_;) (*fn_)(r);
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;)
        {
            uint32_t spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen && !quit_.load(std::memory_order_acquire))
            {
                if (++spins < 20000) { cpu_pause(); continue; }
                std::unique_lock<std::mutex> l(m_);
                cv_.wait_for(l, std::chrono::milliseconds(20), [&]() { return gen_.load(std::memory_order_acquire) != seen || quit_.load(std::memory_order_acquire); });
            }
            if (quit_.load(std::memory_order_acquire)) return;
            seen = gen_.load(std::memory_order_acquire);
            work();
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    unsigned helpers_ = 0;
    std::vector<std::thread> threads_;
    std::mutex m_, call_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<unsigned> pending_{0};
    std::atomic<size_t> next_{0};
    std::atomic<bool> quit_{false};
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t rows_ = 0;
};

void scatter_rows(hb_conv *c, void *const *outs, const char *src, size_t src_ld, size_t src_off, size_t n, int accumulate)
{
    const size_t es = c->esize(), rows_out = host_rows_out(c);
    RowCopyPool::get().run(rows_out, rows_out * n * es, [&](size_t r)
    {
        if (!outs[r]) return;
        const char *sp = src + (r * src_ld + src_off) * es;
        if (!accumulate) memcpy(outs[r], sp, n * es);
        else if (c->dtype == HB_F64)
        {
            double *d = (double *) outs[r];
            const double *q = (const double *) sp;
            for (size_t k = 0; k < n; k++) d[k] += q[k];                   // MonoConvolve.cpp:167-177
        }
        else
        {
            float *d = (float *) outs[r];
            const float *q = (const float *) sp;
            for (size_t k = 0; k < n; k++) d[k] += q[k];
        }
    });
}

void gather_rows(hb_conv *c, const void *const *ins, char *dst, size_t n)
{
    const size_t es = c->esize(), rows_in = host_rows_in(c);
    RowCopyPool::get().run(rows_in, rows_in * n * es, [&](size_t r)
    {
        // a null input row is an inactive channel: silence (NToMonoConvolve.cpp:41 stops at activeInChans)
        if (ins[r]) memcpy(dst + r * n * es, ins[r], n * es);
        else memset(dst + r * n * es, 0, n * es);
    });
}

// Deferred host call, possible when every output sample of this call lies in the block finished by an
// earlier hop (rw + n <= B; the reference reads its output ring before it transforms, PartitionedConvolve.cpp:307
// before :352-360): copy the outputs from the host copy of that block, enqueue this call's work and return
// without waiting for it.  A hop completed by this call is fetched to the other host block behind an event.
// HB_HOST_TRACE=1 (experiments): where a pipelined host call spends its time -- summed per phase, printed when the process exits
struct HostTrace
{
    bool on = false;
    double us[6] = {0, 0, 0, 0, 0, 0};
    uint64_t calls = 0;
    HostTrace() { const char *e = getenv("HB_HOST_TRACE"); on = e && atoi(e); }
    ~HostTrace()
    {
        if (on && calls)
            fprintf(stderr, "[hb host trace] %llu calls, us per call: wait for the pinned input buffer %.1f, gather %.1f, enqueue upload + hop %.1f, "
                            "enqueue download %.1f, wait for the finished block %.1f, scatter %.1f\n", (unsigned long long) calls,
                    us[0] / calls, us[1] / calls, us[2] / calls, us[3] / calls, us[4] / calls, us[5] / calls);
    }
    static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};
HostTrace g_host_trace;

int process_deferred(hb_conv *c, const void *const *ins, void *const *outs, size_t n, int accumulate)
{
    const size_t es = c->esize(), B = c->g.B, rw = c->rw;
    HostTrace &ht = g_host_trace;
    double tm = ht.on ? HostTrace::now() : 0;
    auto lap = [&](int k) { if (ht.on) { const double t = HostTrace::now(); ht.us[k] += t - tm; tm = t; } };
    const size_t rows_in = host_rows_in(c), rows_out = host_rows_out(c);
    int rc;
    for (int k = 0; k < 2; k++)
    {
        if (!c->ev_blk[k]) HB_CUDA(cudaEventCreateWithFlags(&c->ev_blk[k], cudaEventDisableTiming));
        if (!c->ev_inq[k]) HB_CUDA(cudaEventCreateWithFlags(&c->ev_inq[k], cudaEventDisableTiming));
        if (!c->ev_done[k]) HB_CUDA(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
    if (!c->s_h2d) HB_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    if (!c->s_d2h) HB_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    if (c->blk_B != B || c->h_blk[0].cap < rows_out * B * es)
    {
        if ((rc = c->h_blk[0].ensure(rows_out * B * es)) || (rc = c->h_blk[1].ensure(rows_out * B * es))) return rc;
        c->blk_B = B;
        c->blk_valid = false;
        c->blk_pending[0] = c->blk_pending[1] = false;
    }
    if (!c->blk_valid)
    {
        // (re)join the stream: the last completed block lives at the retained head of the staging rows
        if ((rc = drain_deferred(c))) return rc;
        const char *src = (const char *) c->d_yout[c->cur].p + c->y_tail * es;
        HB_CUDA(cudaMemcpy2DAsync(c->h_blk[c->blk_cur].p, B * es, src, c->yout_ld * es, B * es, rows_out, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        c->blk_pending[c->blk_cur] = false;
        c->blk_valid = true;
    }
    // 1. this call's input and device work go out first, so that the GPU is busy while the host copies below
    //    run; the upload and the download of finished blocks ride on their own streams beside the kernels
    const int q = c->inq_cur, have = c->blk_cur;
    if ((rc = c->h_inq[q].ensure(rows_in * n * es)) || (rc = c->d_inq[q].ensure(rows_in * n * es))) return rc;
    if (c->inq_pending[q])
    {
        HB_CUDA(cudaEventSynchronize(c->ev_inq[q]));                        // the pinned buffer has been uploaded
        c->inq_pending[q] = false;
    }
    lap(0);
    gather_rows(c, ins, (char *) c->h_inq[q].p, n);
    lap(1);
    // (forward kernels that read a MiB of rows straight from the pinned buffer themselves took 112 us instead of 12: the rows are
    // brought to the device first)
    if (c->done_pending[q]) HB_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_done[q], 0));   // kernels of two calls ago have read d_inq[q]
    // A MiB-sized upload is done by a copy kernel that reads the pinned rows (mapped host memory) instead of the copy engine: while a
    // tail launch saturates HBM a copy-engine transfer lands only when that launch ends, and the forward FFTs -- and with them the next
    // tail -- start late (profiles/r2_c5_e2e_diag.txt: 106 -> 102 us per call at config 5).  HB_UPLOAD_KERNEL=0: copy engine always.
    static const char *env_upk = getenv("HB_UPLOAD_KERNEL");
    if (!(env_upk && !atoi(env_upk)) && rows_in * n * es >= (size_t(256) << 10))
    {
        void *mapped = nullptr;
        HB_CUDA(cudaHostGetDevicePointer(&mapped, c->h_inq[q].p, 0));
        rc = c->dtype == HB_F64 ? launch_rows<double>((double *) c->d_inq[q].p, n, (const double *) mapped, n, n, rows_in, 0, c->s_h2d)
                                : launch_rows<float>((float *) c->d_inq[q].p, n, (const float *) mapped, n, n, rows_in, 0, c->s_h2d);
        if (rc) return rc;
    }
    else
        HB_CUDA(cudaMemcpyAsync(c->d_inq[q].p, c->h_inq[q].p, rows_in * n * es, cudaMemcpyHostToDevice, c->s_h2d));
    HB_CUDA(cudaEventRecord(c->ev_inq[q], c->s_h2d));
    c->inq_pending[q] = true;
    c->inq_cur = q ^ 1;
    HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_inq[q], 0));
    if ((rc = host_dispatch(c, c->d_inq[q].p, n, nullptr, 0, n, 0, c->stream))) return rc;
    // A MiB-sized block leaves the device as ONE contiguous copy of a packed block (k_rows behind the inverse transforms, a few
    // microseconds): the pitched copy out of the staging rows took about 45 us for 16 rows of 64 KiB, twice the contiguous one, and
    // that copy sits on the round trip block t-1 -> caller -> block t+1 that bounds a host-fed stream (profiles/r2_c5_e2e_diag.txt)
    const int nb = have ^ 1;
    const bool packed = rw + n == B && rows_out * B * es >= (size_t(256) << 10) && c->yout_ld != B;
    if (packed)
    {
        if ((rc = c->d_blk[nb].ensure(rows_out * B * es))) return rc;
        const char *src = (const char *) c->d_yout[c->cur].p + c->y_tail * es;
        rc = c->dtype == HB_F64 ? launch_rows<double>((double *) c->d_blk[nb].p, B, (const double *) src, c->yout_ld, B, rows_out, 0, c->stream)
                                : launch_rows<float>((float *) c->d_blk[nb].p, B, (const float *) src, c->yout_ld, B, rows_out, 0, c->stream);
        if (rc) return rc;
    }
    HB_CUDA(cudaEventRecord(c->ev_done[q], c->stream));
    c->done_pending[q] = true;
    c->blk_valid = true;                    // process_core ran no reset here: ensure_ready was called by the caller
    lap(2);
    if (rw + n == B)
    {
        const char *src = (const char *) c->d_yout[c->cur].p + c->y_tail * es;
        HB_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_done[q], 0));
        if (packed) HB_CUDA(cudaMemcpyAsync(c->h_blk[nb].p, c->d_blk[nb].p, rows_out * B * es, cudaMemcpyDeviceToHost, c->s_d2h));
        else HB_CUDA(cudaMemcpy2DAsync(c->h_blk[nb].p, B * es, src, c->yout_ld * es, B * es, rows_out, cudaMemcpyDeviceToHost, c->s_d2h));
        HB_CUDA(cudaEventRecord(c->ev_blk[nb], c->s_d2h));
        c->blk_pending[nb] = true;
        c->blk_cur = nb;
    }
    lap(3);
    // 2. the outputs of this call: samples rw .. rw+n of the block an earlier hop finished
    if (c->blk_pending[have])
    {
        HB_CUDA(cudaEventSynchronize(c->ev_blk[have]));
        c->blk_pending[have] = false;
        // that block came out of the hop of the previous call, whose kernels had read the upload of that call: the other pinned input
        // buffer is free without asking its event (a cudaEventSynchronize on a finished event still costs several microseconds)
        c->inq_pending[q ^ 1] = false;
    }
    lap(4);
    scatter_rows(c, outs, (const char *) c->h_blk[have].p, B, rw, n, accumulate);
    lap(5);
    if (ht.on) ht.calls++;
    return HB_OK;
}
} // namespace

extern "C" int hb_conv_process(hb_conv *c, const void *const *ins, void *const *outs, uintptr_t n, int accumulate)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if ((!ins || !outs) && n) { set_error("hb_conv_process: null buffer"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    // (a rank of the fused multi-GPU exchange takes part in every hop, loaded or not: its peers wait for its blocks)
    if (!c->P && !c->peers_attached) return HB_ERR_NO_IR;
    if (!n) return HB_OK;
    rc = c->dtype == HB_F64 ? ensure_ready<double>(c, n, c->stream) : ensure_ready<float>(c, n, c->stream);
    if (rc) return rc;
    if (c->deferred && c->rw + n <= c->g.B) return process_deferred(c, ins, outs, n, accumulate);

    // a call that crosses a hop boundary needs that hop's result before it returns: synchronous round trip
    const size_t es = c->esize();
    const size_t rows_in = host_rows_in(c), rows_out = host_rows_out(c);
    if ((rc = c->h_in.ensure(rows_in * n * es)) || (rc = c->h_out.ensure(rows_out * n * es)) ||
        (rc = c->d_io_in.ensure(rows_in * n * es)) || (rc = c->d_io_out.ensure(rows_out * n * es))) return rc;
    HB_CUDA(cudaStreamSynchronize(c->stream));                 // deferred work may still be reading the staging buffers
    if ((rc = drain_deferred(c))) return rc;
    gather_rows(c, ins, (char *) c->h_in.p, n);
    HB_CUDA(cudaMemcpyAsync(c->d_io_in.p, c->h_in.p, rows_in * n * es, cudaMemcpyHostToDevice, c->stream));
    if ((rc = host_dispatch(c, c->d_io_in.p, n, c->d_io_out.p, n, n, 0, c->stream))) return rc;
    HB_CUDA(cudaMemcpyAsync(c->h_out.p, c->d_io_out.p, rows_out * n * es, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    c->blk_valid = false;
    scatter_rows(c, outs, (const char *) c->h_out.p, n, 0, n, accumulate);
    return HB_OK;
}

extern "C" int hb_conv_shard_export(hb_conv *c, uint32_t world, uint32_t rank, void *handle_out)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (world < 1 || world > HB_MAX_WORLD || rank >= world || c->groups != 1 || c->outs % world)
    {
        set_error("hb_conv_shard_export: needs 1 <= world <= %d, rank < world, one group and outs a multiple of world", HB_MAX_WORLD);
        return HB_ERR_BAD_ARG;
    }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (c->d_inbox) { set_error("hb_conv_shard_export: already exported"); return HB_ERR_BAD_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == HB_IPC_HANDLE_BYTES, "IPC handle size");
    const size_t slot = (size_t(1) << c->max_fft_log2) >> 1;
    const size_t data = (size_t(HB_INBOX_DEPTH) * world * (c->outs / world) * slot * c->esize() + 255) & ~size_t(255);
    HB_CUDA(cudaMalloc(&c->d_inbox, data + INBOX_TAIL_BYTES));
    HB_CUDA(cudaMemset(c->d_inbox, 0, data + INBOX_TAIL_BYTES));
    HB_CUDA(cudaDeviceSynchronize());
    if (handle_out)                 // NULL: the peers live in this process (hb_conv_shard_attach_local)
    {
        cudaIpcMemHandle_t h;
        HB_CUDA(cudaIpcGetMemHandle(&h, c->d_inbox));
        memcpy(handle_out, &h, sizeof(h));
    }
    c->shard_world = world; c->shard_rank = rank;
    c->inbox_data_bytes = data; c->inbox_slot = slot;
    c->hop_seq = 0;
    for (uint32_t k = 0; k < HB_INBOX_DEPTH; k++) c->parity_uses[k] = 0;
    return HB_OK;
}

extern "C" int hb_conv_shard_attach(hb_conv *c, const void *handles)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (!handles || !c->d_inbox || c->peers_attached) { set_error("hb_conv_shard_attach: export first, attach once"); return HB_ERR_BAD_ARG; }
    for (uint32_t r = 0; r < c->shard_world; r++)
    {
        if (r == c->shard_rank) { c->peer_base[r] = c->d_inbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + size_t(r) * HB_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
        {
            set_error("cudaIpcOpenMemHandle for rank %u -> %s (peer access over NVLink / PCIe is required)", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (uint32_t k = 0; k < r; k++)
                if (k != c->shard_rank && c->peer_base[k]) { cudaIpcCloseMemHandle(c->peer_base[k]); c->peer_base[k] = nullptr; }
            return HB_ERR_CUDA;
        }
        c->peer_base[r] = p;
    }
    c->peers_attached = true;
    return HB_OK;
}

extern "C" int hb_conv_shard_attach_local(hb_conv *c, hb_conv *const *peers)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (!peers || !c->d_inbox || c->peers_attached) { set_error("hb_conv_shard_attach_local: export first, attach once"); return HB_ERR_BAD_ARG; }
    for (uint32_t r = 0; r < c->shard_world; r++)
    {
        if (r == c->shard_rank) { c->peer_base[r] = c->d_inbox; continue; }
        const hb_conv *p = peers[r];
        if (!p || !p->d_inbox || p->shard_world != c->shard_world || p->shard_rank != r || p->inbox_data_bytes != c->inbox_data_bytes)
        {
            set_error("hb_conv_shard_attach_local: peer %u is not an exported engine of the same exchange", r);
            return HB_ERR_BAD_ARG;
        }
        if (p->device != c->device)
        {
            int can = 0;
            HB_CUDA(cudaDeviceCanAccessPeer(&can, c->device, p->device));
            if (!can) { set_error("device %d cannot map the memory of device %d (peer access over NVLink / PCIe is required)", c->device, p->device); return HB_ERR_UNSUPPORTED; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d) -> %s", p->device, cudaGetErrorString(e)); cudaGetLastError(); return HB_ERR_CUDA; }
        }
        c->peer_base[r] = p->d_inbox;
    }
    c->peers_attached = true;
    c->peers_local = true;
    return HB_OK;
}

extern "C" int hb_conv_shard_status(hb_conv *c, uint32_t *late_ranks)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (!c->d_inbox || !late_ranks) { set_error("hb_conv_shard_status: not a sharded engine"); return HB_ERR_BAD_ARG; }
    HB_CUDA(cudaDeviceSynchronize());
    HB_CUDA(cudaMemcpy(late_ranks, (const char *) c->d_inbox + c->inbox_data_bytes + INBOX_LATE_OFF, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_conv_process_shard_dev(hb_conv *c, const void *d_in, uintptr_t in_ld, void *d_out_shard, uintptr_t out_ld,
                                         uintptr_t num_samples, int accumulate, void *stream)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if ((!d_in || !d_out_shard) && num_samples) { set_error("hb_conv_process_shard_dev: null buffer"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    if (!c->peers_attached) { set_error("hb_conv_process_shard_dev: peers not attached"); return HB_ERR_BAD_ARG; }
    cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
    if (c->blk_valid && (rc = drain_deferred(c))) return rc;
    c->blk_valid = false;
    return shard_dispatch(c, d_in, in_ld, d_out_shard, out_ld, num_samples, accumulate, st);
}

extern "C" int hb_conv_join(hb_conv *c, void *stream)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
    // the only work an engine keeps in flight outside the caller's stream: the tail launched ahead for the next hop
    if (c->split && c->tail_valid) HB_CUDA(cudaStreamWaitEvent(st, c->ev_tail[c->tail_par], 0));
    return HB_OK;
}

extern "C" int hb_conv_set_hop_overlap(hb_conv *c, int mode)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (mode < 0 || mode > 2) { set_error("hop overlap mode must be 0 (never), 1 (the engine's own stream) or 2 (any stream)"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->hop_overlap = mode;
    return HB_OK;
}

extern "C" int hb_conv_set_tuning(hb_conv *c, int ctas_per_sm, int variant)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (variant != 0 && variant != 1) { set_error("variant must be 0 (direct loads) or 1 (TMA ring)"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->ctas_per_sm = ctas_per_sm > 0 ? ctas_per_sm : (variant == 1 ? 1 : 2);
    c->variant = variant;
    c->need_reset = true;           // partial-segment geometry depends on the grid
    return HB_OK;
}

extern "C" int hb_conv_set_tail_streams(hb_conv *c, int streams)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (streams < 0 || streams > 2) { set_error("tail streams must be 0 (automatic), 1 or 2"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->tail_streams_req = streams;
    c->need_reset = true;
    return HB_OK;
}

extern "C" int hb_conv_set_schedule(hb_conv *c, int overlapped)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (overlapped < 0 || overlapped > 3) { set_error("schedule must be 0 (serial), 1 (overlapped), 2 (automatic) or 3 (fused where eligible)"); return HB_ERR_BAD_ARG; }
    c->schedule = overlapped;
    c->need_reset = true;           // the partial-segment sets depend on the schedule
    return HB_OK;
}

extern "C" int hb_conv_set_fft_path(hb_conv *c, int path)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (path < 0 || path > 3) { set_error("fft path must be 0 (automatic), 1 (one CTA per transform), 2 (cluster of 8 CTAs) or 3 (four-step)"); return HB_ERR_BAD_ARG; }
    c->fft_path = path;
    c->need_reset = true;           // the tail grid depends on where the FFT CTAs run
    return HB_OK;
}

extern "C" int hb_conv_fft_path(const hb_conv *c)
{
    if (!c) return -1;
    const int m = (int) c->g.log2n - 1;
    if (m > single_cta_max_log2m(c)) return 3;
    return use_cluster_fft(c) ? 2 : 1;
}

extern "C" int hb_conv_set_multi_hop(hb_conv *c, int enable)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    c->multi_hop = enable ? 1 : 0;
    c->need_reset = true;
    return HB_OK;
}

extern "C" int hb_conv_set_trace(hb_conv *c, int enable)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    HB_CUDA(cudaDeviceSynchronize());
    if (!enable) { c->d_trace.release(); return HB_OK; }
    const size_t bytes = size_t(TRACE_HOPS) * TRACE_KINDS * 2 * TRACE_CTAS * sizeof(unsigned long long);
    if ((rc = c->d_trace.ensure(bytes))) return rc;
    HB_CUDA(cudaMemset(c->d_trace.p, 0, bytes));
    return HB_OK;
}

extern "C" int hb_conv_get_trace(hb_conv *c, uint64_t *out, uint64_t *hop)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if (!c->d_trace.p || !out) { set_error("hb_conv_get_trace: tracing is off"); return HB_ERR_BAD_ARG; }
    HB_CUDA(cudaDeviceSynchronize());
    HB_CUDA(cudaMemcpy(out, c->d_trace.p, size_t(TRACE_HOPS) * TRACE_KINDS * 2 * TRACE_CTAS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (hop) *hop = c->g.hop;
    return HB_OK;
}

extern "C" int hb_conv_set_host_pipeline(hb_conv *c, int pipelined)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if ((rc = drain_deferred(c))) return rc;
    c->blk_valid = false;
    c->deferred = pipelined != 0;
    return HB_OK;
}

extern "C" int hb_conv_set_profiling(hb_conv *c, int enable)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if ((rc = drain_profile(c))) return rc;
    c->profiling = enable != 0;
    for (int k = 0; k < 5; k++) c->prof_ms[k] = 0;
    c->prof_hops = 0;
    return HB_OK;
}

extern "C" int hb_conv_get_profile(hb_conv *c, double *ms, uint64_t *hops)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    c->chain_ok = false;
    if ((rc = drain_profile(c))) return rc;
    if (ms) for (int k = 0; k < 5; k++) ms[k] = c->prof_ms[k];
    if (hops) *hops = c->prof_hops;
    return HB_OK;
}

extern "C" int hb_conv_tail_streams(const hb_conv *c) { return c ? c->tail_streams : 0; }
extern "C" int hb_conv_schedule(const hb_conv *c) { return !c ? 0 : (c->fused ? 2 : (c->split ? 1 : 0)); }

extern "C" uint64_t hb_conv_bytes_per_launch(const hb_conv *c)
{
    if (!c || !c->P) return 0;
    // the dominant multiply-accumulate launch: the whole hop in the serial schedule (hb_conv_bytes_per_hop), in the
    // overlapped one the tail's IR partitions 1..P-1 and the P-1 spectra of every input they meet
    if (!c->split) return hb_conv_bytes_per_hop(c);
    const uint64_t s = c->esize(), B = (uint64_t(1) << c->fft_log2) >> 1, P = c->P;
    const uint64_t I = uint64_t(c->groups) * c->ins, K = c->pairs();
    return 2 * s * B * (P - 1) * (K + I);
}

extern "C" uint64_t hb_conv_bytes_per_hop(const hb_conv *c)
{
    if (!c || !c->P) return 0;
    // 2sB*P*K (IR spectra) + 2sB*P*I (FDL) + sB*(I+O), K = pairs, per group
    const uint64_t s = c->esize(), B = (uint64_t(1) << c->fft_log2) >> 1, P = c->P;
    const uint64_t I = uint64_t(c->groups) * c->ins, O = uint64_t(c->groups) * c->outs, K = c->pairs();
    return 2 * s * B * P * K + 2 * s * B * P * I + s * B * (I + O);
}
