/* include/hisstools_b200.h -- the drop-in boundary: C ABI of libhisstools_b200.so
 *
 * B200-native (sm_100a) implementation of the HISSTools_Library partitioned-convolution hot path.
 * Plain pointers and sizes only; no C++ or torch types.  The C++ headers under
 * include/HISSTools_FFT and include/HIRT_Multichannel_Convolution are thin source-compatible
 * wrappers over these entry points, and hisstools_library_b200/ (python, ctypes) binds the same
 * symbols for tests and bench.py.  There is no CPU fallback anywhere behind this boundary: every
 * compute entry point returns HB_ERR_CUDA when no device is usable.
 *
 * Each entry point cites the reference interface it replaces (paths under the reference tree).
 *
 * Conventions
 *   - return value: a reference ConvolveError code 0..12 (HIRT_Multichannel_Convolution/ConvolveErrors.h:4-19)
 *     or a negative hb_status for conditions the reference cannot express.
 *   - dtype: HB_F32 (0) or HB_F64 (1).  The reference convolver classes are float-only
 *     (PartitionedConvolve.h:38-41); HB_F64 is the double engine BASELINE config 5 asks for.
 *   - host pointers are only read/written during the call (PartitionedConvolve.cpp:212-217);
 *     `_dev` variants take device pointers on the handle's device and enqueue on `stream`
 *     (a cudaStream_t passed as void*; NULL = the handle's own stream) without synchronising.
 */
#ifndef HISSTOOLS_B200_H
#define HISSTOOLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_F32 0
#define HB_F64 1

typedef enum hb_status
{
    HB_OK = 0,
    HB_ERR_CUDA = -1,          /* no device / CUDA runtime failure (see hb_last_error) */
    HB_ERR_BAD_ARG = -2,
    HB_ERR_UNSUPPORTED = -3,   /* size outside what the library implements */
    HB_ERR_NO_IR = -4,         /* process() on an object with no impulse response: outputs untouched */
    HB_ERR_BUSY = -5           /* process() while set()/resize() holds the object: block skipped, outputs untouched */
} hb_status;

/* last CUDA / argument error text of the calling thread ("" if none) */
const char *hb_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t hb_launch_count(void);
/* library build identification, e.g. "hisstools_b200 sm_100a" */
const char *hb_version(void);

/* ---------------------------------------------------------------------------------------------
 * FFT family -- replaces HISSTools_FFT/HISSTools_FFT.h:87-369 (hisstools_create_setup,
 * hisstools_destroy_setup, hisstools_fft / ifft / rfft / rifft, and the out-of-place real
 * conveniences).  Same conventions: split-complex planes, forward kernel exp(-j theta), real
 * forward transform returns 2*DFT with DC in realp[0] and Nyquist in imagp[0], nothing is scaled
 * (rifft(rfft(x)) = 2N x).  Host pointers; each call is one H2D, one kernel, one D2H.
 * ------------------------------------------------------------------------------------------- */
typedef struct hb_fft_setup hb_fft_setup;

/* hisstools_create_setup(FFT_SETUP_F/D*, max_fft_log_2): HISSTools_FFT.h:87,98 */
int hb_fft_setup_create(hb_fft_setup **out, int dtype, uintptr_t max_fft_log2, int device);
/* hisstools_destroy_setup: HISSTools_FFT.h:108,118 */
void hb_fft_setup_destroy(hb_fft_setup *setup);

/* in-place complex transforms on split planes of 2^log2n points: HISSTools_FFT.h:130,142 (fft) 220,232 (ifft) */
int hb_fft(hb_fft_setup *setup, void *realp, void *imagp, uintptr_t log2n);
int hb_ifft(hb_fft_setup *setup, void *realp, void *imagp, uintptr_t log2n);
/* in-place real transforms; planes hold 2^(log2n-1) points: HISSTools_FFT.h:154,166 (rfft) 244,256 (rifft) */
int hb_rfft(hb_fft_setup *setup, void *realp, void *imagp, uintptr_t log2n);
int hb_rifft(hb_fft_setup *setup, void *realp, void *imagp, uintptr_t log2n);
/* out-of-place real forward with zero padding: HISSTools_FFT.h:180,194,208.
 * in_dtype may be HB_F32 with an HB_F64 setup (the float->double overload, :208). */
int hb_rfft_real(hb_fft_setup *setup, const void *input, int in_dtype, void *realp, void *imagp,
                 uintptr_t in_length, uintptr_t log2n);
/* out-of-place real inverse (planes are left holding the de-interleaved result, as the reference): HISSTools_FFT.h:269,282 */
int hb_rifft_real(hb_fft_setup *setup, void *realp, void *imagp, void *output, uintptr_t log2n);

/* batched device-pointer variants (benchmarking / pipelines without PCIe).  `batch` transforms,
 * consecutive transforms `stride` elements apart in every array. */
int hb_rfft_real_batched_dev(hb_fft_setup *setup, const void *d_input, void *d_realp, void *d_imagp,
                             uintptr_t log2n, uintptr_t batch, uintptr_t in_stride, uintptr_t out_stride, void *stream);
int hb_rifft_real_batched_dev(hb_fft_setup *setup, const void *d_realp, const void *d_imagp, void *d_output,
                              uintptr_t log2n, uintptr_t batch, uintptr_t in_stride, uintptr_t out_stride, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Uniform partitioned convolution engine.  One handle = `groups` independent banks, each an
 * `ins` x `outs` matrix of impulse responses convolved by overlap-save with hop B = fft_size/2:
 *     out[g][o] = sum_i  IR[g][o][i] * in[g][i]        (delayed by exactly B samples)
 *   groups=1, ins=outs=1     HISSTools::PartitionedConvolve  (PartitionedConvolve.h:23-41)
 *   groups=1, ins=N, outs=1  HISSTools::NToMonoConvolve::process (NToMonoConvolve.cpp:35-43), uniform partitions
 *   groups=1, ins=N, outs=M  HISSTools::Convolver, N x M mode (Convolver.cpp:5-22,138-154)
 *   groups=K, ins=outs=1     HISSTools::Convolver, parallel mode (Convolver.cpp:24-41)
 * Per hop and bank: `ins` forward real FFTs, one frequency-domain multiply-accumulate over
 * (input x partition) for every output, `outs` inverse FFTs, scale 1/(4N), first B samples kept
 * (PartitionedConvolve.cpp:352-377 with the input-sum of NToMonoConvolve.cpp:39-42 moved into
 * the frequency domain).
 * ------------------------------------------------------------------------------------------- */
typedef struct hb_conv hb_conv;

/* PartitionedConvolve(maxFFTSize, maxLength, offset, length): PartitionedConvolve.cpp:52-102.
 * max_length is rounded up to a multiple of max_fft_size/2 as the reference does (:77-82).
 * Returns the constructor-time error of setMaxFFTSize (:26-50), which the reference discards;
 * the handle is created (with the clamped size) in every case except HB_ERR_*. */
int hb_conv_create(hb_conv **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs,
                   uintptr_t max_fft_size, uintptr_t max_length, uintptr_t offset, uintptr_t length, int device);
void hb_conv_destroy(hb_conv *c);

/* setFFTSize / setLength / setOffset / setResetOffset: PartitionedConvolve.cpp:131-171.
 * A changed FFT size drops every loaded IR until hb_conv_set_ir is called again (:145-149).
 * A negative reset offset (the reference's "random phase", :275-278) selects phase 0 here: the
 * result is phase independent up to rounding and a fixed phase keeps runs reproducible. */
int hb_conv_set_fft_size(hb_conv *c, uintptr_t fft_size);
int hb_conv_set_length(hb_conv *c, uintptr_t length);
int hb_conv_set_offset(hb_conv *c, uintptr_t offset);
int hb_conv_set_reset_offset(hb_conv *c, intptr_t offset);

/* set(input, length): PartitionedConvolve.cpp:173-225 for pair (group, in, out); ir_dtype is the
 * element type of `ir` (HB_F32 or HB_F64; converted to the engine's type as Convolver.cpp:126-134
 * does for double IRs).  ir == NULL or length <= offset clears the pair.  Triggers reset(). */
int hb_conv_set_ir(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length);
/* set() on an engine that is running, restarting THIS pair only -- what the reference does (MonoConvolve::set resets one object,
 * MonoConvolve.cpp:118-140, while the other pairs of a Convolver keep playing).  The pair's spectra are hidden and come back one
 * partition per hop, partition p at the hop where the frame it meets is the first one recorded after the call, so the new response
 * starts from silence and never meets earlier input (up to the fft_size / 2 samples before the call that share its first frame;
 * the block finished before the call is still delivered with the old response).  While a pair is coming back, hops run one by
 * one with all partitions in one launch.  Falls back to hb_conv_set_ir (whole stream restarts) when the stream is not running yet,
 * the new response needs a longer delay line than the engine has, or the FFT size is above the one-CTA limit. */
int hb_conv_set_ir_live(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length);
/* reset of one pair (Convolver::reset(in, out), Convolver.cpp:88-97): the pair forgets its input history as above, keeping its response;
 * on an engine that is not running, or where it cannot be done in place, the whole stream restarts (hb_conv_reset) */
int hb_conv_reset_pair(hb_conv *c, uint32_t group, uint32_t in, uint32_t out);
/* same, `d_ir` in device memory in the engine's dtype.  Synchronises: work enqueued earlier on ANY stream of the device
 * (whatever produced d_ir) is waited for before the transforms start, and the call returns when the spectra are in place
 * (d_ir may be reused at once). */
int hb_conv_set_ir_dev(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *d_ir, uintptr_t length);
/* grow / shrink the allocation to hold max_length taps per pair (the MemorySwap::equal step of
 * MonoConvolve.cpp:100-110).  Loaded IRs are kept (cut to the new capacity when it shrinks); the
 * stream restarts from silence at the next process call. */
int hb_conv_resize(hb_conv *c, uintptr_t max_length);
/* reset(): PartitionedConvolve.cpp:227-230 -- takes effect at the next process call */
int hb_conv_reset(hb_conv *c);

/* number of partitions currently loaded (0 = no IR: process returns HB_ERR_NO_IR and leaves outputs untouched) */
uintptr_t hb_conv_partitions(const hb_conv *c);
uintptr_t hb_conv_max_length(const hb_conv *c);
uintptr_t hb_conv_fft_size(const hb_conv *c);

/* process(in, out, numSamples): PartitionedConvolve.cpp:243-385 for every bank at once.
 * ins: groups*ins planar host pointers (bank-major), outs: groups*outs planar host pointers, any
 * numSamples (state is carried across calls, :257,298-299,382).  accumulate != 0 adds into outs
 * (MonoConvolve.cpp:167-177) instead of overwriting. */
int hb_conv_process(hb_conv *c, const void *const *ins, void *const *outs, uintptr_t num_samples, int accumulate);
/* device-resident variant: d_in is [groups*ins][in_ld] and d_out [groups*outs][out_ld] in the
 * engine's dtype; enqueued on `stream`, no synchronisation. */
int hb_conv_process_dev(hb_conv *c, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld,
                        uintptr_t num_samples, int accumulate, void *stream);

/* Multi-GPU: the N x M matrix with its INPUT channels sharded over `world` ranks (one process per GPU).
 * The sum over inputs of NToMonoConvolve.cpp:39-42 then crosses devices; it is fused into the inverse-FFT
 * epilogue: every rank stores each partial output block straight into the inbox of the rank that owns
 * that output (peer-mapped memory, NVLink) and the owner sums the arrivals in rank order.
 *   export: allocate this rank's inbox and return its CUDA IPC handle (HB_IPC_HANDLE_BYTES bytes);
 *   attach: after the callers exchanged the handles (any transport), open all of them (rank order);
 *   process_shard_dev: one hop-aligned call (num_samples a multiple of fft_size/2, reset offset 0) with
 *   this rank's `ins` input rows; d_out_shard receives this rank's outs/world output rows (outputs
 *   rank*outs/world ...), complete sums, delayed by fft_size/2 as ever.  Every rank must make the same
 *   calls; reset / set_ir are collective in that sense. */
#define HB_IPC_HANDLE_BYTES 64
int hb_conv_shard_export(hb_conv *c, uint32_t world, uint32_t rank, void *handle_out);
int hb_conv_shard_attach(hb_conv *c, const void *handles);
int hb_conv_process_shard_dev(hb_conv *c, const void *d_in, uintptr_t in_ld, void *d_out_shard, uintptr_t out_ld,
                              uintptr_t num_samples, int accumulate, void *stream);

/* The same exchange between engines that live in ONE process (one per GPU): hb_conv_shard_export with handle_out = NULL on every
 * engine, then attach_local on every engine with the array of all `world` engine handles in rank order.  Peer access between the
 * devices is enabled here (cudaDeviceEnablePeerAccess).  hb_matrix_create_multi does this for a whole matrix. */
int hb_conv_shard_attach_local(hb_conv *c, hb_conv *const *peers);
/* late_ranks: bit r set = the owner-side sum on this rank gave up waiting for rank r's blocks at least once (longer than
 * HB_PEER_TIMEOUT_MS, default 30 s) and summed that hop without them.  Synchronises the device.  A rank that has no
 * impulse response loaded still takes part in every hop (it delivers silence), so only a stopped peer raises this. */
int hb_conv_shard_status(hb_conv *c, uint32_t *late_ranks);

/* Make `stream` (NULL = the handle's own) wait for the work the engine keeps in flight on its internal streams -- the
 * share of the NEXT hop that the overlapped schedule launches ahead (hb_conv_set_schedule).  Callers that time a run of
 * process_dev calls with events on their stream call this before the closing event, so that the timed window holds as
 * many tail launches as hops. */
int hb_conv_join(hb_conv *c, void *stream);

/* tuning / introspection used by bench.py and the tests */
/* CTAs per SM for the multiply-accumulate kernel (0 = library default) */
int hb_conv_set_tuning(hb_conv *c, int ctas_per_sm, int variant);
/* host-pointer calls (hb_conv_process) that stay inside one hop are pipelined by default: the call returns the
 * samples an earlier hop finished (the API's own latency of fft_size/2, PartitionedConvolve.cpp:307 before
 * :352-360) and leaves its own device work in flight.  pipelined = 0 makes every call a synchronous round trip. */
int hb_conv_set_host_pipeline(hb_conv *c, int pipelined);
/* algorithmic bytes one hop moves (IR spectra + FDL + time-domain I/O) */
uint64_t hb_conv_bytes_per_hop(const hb_conv *c);
/* Schedule of a hop.  overlapped = 1: the share of the NEXT hop that only needs spectra already in the
 * delay line (partitions 1..P-1, "tail") is computed on a second stream as soon as this hop's forward FFTs are
 * done, beside the inverse FFTs of this hop and the forward FFTs of the next; a hop's critical path is forward
 * FFT -> partition 0 ("head") -> inverse FFT.  It is the stream form of the reference's spreading of partitions
 * over the samples of a hop (PartitionedConvolve.cpp:330-347).  overlapped = 0: forward FFTs, one multiply-accumulate
 * over all partitions, inverse FFTs, in a row.  overlapped = 2 (default, automatic): the fused hop where it applies,
 * else overlapped when the tail streams at least 4 MiB of spectra per hop, else serial (launch-latency-bound hops gain
 * nothing from a second stream).  overlapped = 3: as 2.  The FUSED hop is one thread-block-cluster launch per hop for
 * single-output engines with small spectra (PartitionedConvolve, MonoConvolve parts, NToMonoConvolve): the ranks of a
 * cluster transform the inputs, split the (input, partition) products and reduce through distributed shared memory
 * (hb_conv_fused.cuh).  overlapped = 0 / 1 never fuse.  Results differ by summation order only.  Takes effect with a reset. */
int hb_conv_set_schedule(hb_conv *c, int overlapped);
/* Overlapped schedule only: 1 = every tail launch on one second stream, one after the other; 2 = the tails of consecutive hops
 * alternate between two streams, so the CTAs of the next tail are placed on the SMs as the CTAs of the running one leave (hides
 * part of the ramp-up and drain of a launch: profiles/r2_tail_streams.txt); 0 (default) = automatic: 2 when a tail launch streams
 * 3 GiB or less.  Takes effect with a reset.  hb_conv_tail_streams: the number in effect. */
int hb_conv_set_tail_streams(hb_conv *c, int streams);
int hb_conv_tail_streams(const hb_conv *c);
/* Fused hops only: may consecutive single-hop calls run side by side?  A hop needs the block the previous call saved and the
 * spectra of earlier hops, not the previous hop's output, so hop t+1 can transform its frame while hop t still multiplies -- if its
 * input rows may be read before the stream has finished what precedes the launch.  mode 1 (default): only for calls on the engine's
 * own stream (`stream` = NULL), where rows cannot be ordered behind the caller's work anyway and must be complete when the call is
 * made; mode 2: on any stream -- the caller declares that the input rows of a call are complete when it is made and that nothing
 * the engine has to wait for is enqueued between two calls; mode 0: never.  Outputs are written in stream order in every mode and
 * the samples are the same; only the time between back-to-back calls changes (profiles/r2_small_hops.txt). */
int hb_conv_set_hop_overlap(hb_conv *c, int mode);
/* schedule in effect after the last reset: 0 serial (also whenever only one partition is loaded), 1 overlapped, 2 fused */
int hb_conv_schedule(const hb_conv *c);
/* algorithmic bytes of the dominant multiply-accumulate launch: hb_conv_bytes_per_hop in the serial schedule; in
 * the overlapped one the tail's share, 2sB(P-1)(K+I) */
uint64_t hb_conv_bytes_per_launch(const hb_conv *c);
/* Multi-hop reuse (default on): a hop-aligned call that brings several hops at once (numSamples >= 2 * fft_size / 2) has all
 * their input spectra in the delay line before any output is needed, so HBM-bound engines stream every impulse-response
 * spectrum ONCE for up to four hops (k_cmac_tma_mh) instead of once per hop.  Same result up to summation order;
 * enable = 0 processes such calls hop by hop.  Takes effect with a reset. */
int hb_conv_set_multi_hop(hb_conv *c, int enable);
/* Where the forward / inverse transforms of a hop run.  0 (default, automatic), 1: one CTA per channel (k_fwd / k_inv);
 * 2: every transform spread over a thread-block cluster of 8 CTAs exchanging through distributed shared memory
 * (hb_conv_cluster.cuh; transforms of 2^11 points up to the one-CTA limit, single hops, not the fused multi-GPU exchange --
 * otherwise the one-CTA kernels run); 3: the four-step chains over global memory (hb_conv_big.cuh) from 2^12 points.
 * Automatic = 2 for double-precision engines with fewer transforms than two per SM, whose one-CTA transforms (long,
 * latency-bound, on a few SMs) would otherwise set the hop period; sizes above the one-CTA limit always take the four-step chains.
 * Results differ by rounding only.  Takes effect with a reset.  hb_conv_fft_path: the path in effect (1, 2 or 3). */
int hb_conv_set_fft_path(hb_conv *c, int path);
int hb_conv_fft_path(const hb_conv *c);
/* Kernel timeline (debug): while enabled, thread 0 of every CTA of the hop kernels stamps %globaltimer at entry
 * and exit.  hb_conv_get_trace synchronises the device and copies the HB_TRACE_WORDS stamps out:
 * out[(((hop % 16) * 5 + kind) * 2 + exit) * 256 + cta], kind 0 forward FFT, 1 head, 2 tail / whole
 * multiply-accumulate, 3 inverse FFT, 4 owner-side sum of the multi-GPU exchange; *hop = hops processed so far
 * (tools/trace_timeline.py prints it). */
#define HB_TRACE_WORDS (16 * 5 * 2 * 256)
int hb_conv_set_trace(hb_conv *c, int enable);
int hb_conv_get_trace(hb_conv *c, uint64_t *out, uint64_t *hop);
/* per-kernel device timing: while enabled, every hop records CUDA events around its kernels on the streams they
 * are launched on.  hb_conv_get_profile waits for the recorded hops and returns in ms[0..4] the summed milliseconds
 * of: forward FFTs; the whole (serial) or head (overlapped) multiply-accumulate; the wait for the tail; inverse
 * FFTs; the tail multiply-accumulate (0 when serial) -- and the hop count since profiling was enabled.  (An event
 * recorded straight behind a stream wait is not ordered after it: only ms[2] + ms[3] is meaningful when overlapped.) */
int hb_conv_set_profiling(hb_conv *c, int enable);
int hb_conv_get_profile(hb_conv *c, double *ms, uint64_t *hops);

/* ---------------------------------------------------------------------------------------------
 * Non-uniform partition scheme for a whole channel matrix -- what MonoConvolve (MonoConvolve.h:30-48)
 * is for one pair, for `groups` banks of ins x outs pairs at once.  This is the object the C++
 * classes MonoConvolve (1x1), NToMonoConvolve (N x 1) and Convolver (N x M, or K parallel banks of
 * 1x1) under include/HIRT_Multichannel_Convolution forward to.
 *   scheme: a direct-form zero-latency head of A/2 taps when zero_latency (TimeDomainConvolve.cpp:69-163),
 *   fixed parts of FFT size A, B, C covering (next - size)/2 taps each and a resizable tail of the
 *   largest size (MonoConvolve.cpp:203-258).  Net delay: 0 with zero_latency, A/2 otherwise.
 * ------------------------------------------------------------------------------------------- */
typedef struct hb_matrix hb_matrix;

/* MonoConvolve(maxLength, zeroLatency, A, B, C, D): MonoConvolve.cpp:36-45.  An invalid size list
 * (the reference throws std::runtime_error, :212,229) returns HB_ERR_BAD_ARG. */
int hb_matrix_create(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                     int zero_latency, uint32_t A, uint32_t B, uint32_t C, uint32_t D, int device);
/* MonoConvolve(maxLength, LatencyMode): MonoConvolve.cpp:18-32; latency_mode 0 zero, 1 short, 2 medium (MonoConvolve.h:14-19) */
int hb_matrix_create_latency(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                             int latency_mode, int device);
void hb_matrix_destroy(hb_matrix *m);
/* MonoConvolve::setResetOffset: MonoConvolve.cpp:80-98 (negative selects phase 0).  The reference staggers its parts by
 * size / 8 samples to spread CPU load over callbacks; here every part takes the offset as it is (the output does not depend
 * on the phase beyond rounding, and hop-aligned calls are what the GPU path is fastest on). */
int hb_matrix_set_reset_offset(hb_matrix *m, intptr_t offset);
/* MonoConvolve::resize / set / reset for pair (group, in, out): MonoConvolve.cpp:100-152.  Return the
 * reference ConvolveError (0, 3, 4) or a negative hb_status.  set/resize block process (MemorySwap.h:174-178). */
int hb_matrix_resize(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out, uintptr_t length);
int hb_matrix_set(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length, int request_resize);
int hb_matrix_reset(hb_matrix *m);
/* Convolver::reset(inChan, outChan) (Convolver.cpp:88-97): restarts that pair only (every part and the head; hb_conv_reset_pair).
 * hb_matrix_set / hb_matrix_resize on a running matrix likewise restart only the pair they change (hb_conv_set_ir_live). */
int hb_matrix_reset_pair(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out);
/* hb_conv_set_hop_overlap for the parts of a matrix, applied to hb_matrix_process_dev calls: 1 (default) = calls on the matrix's
 * own stream (`stream` = NULL), 2 = any stream (input rows complete when the call is made), 0 = never. */
int hb_matrix_set_hop_overlap(hb_matrix *m, int mode);
/* MonoConvolve::process / NToMonoConvolve::process / Convolver::process (MonoConvolve.cpp:179-201,
 * NToMonoConvolve.cpp:35-43, Convolver.cpp:138-154): ins = groups*ins planar host rows (NULL row =
 * inactive input, silence), outs = groups*outs planar host rows (NULL row = not wanted).
 * HB_OK: outs written (accumulate != 0: added to).  HB_ERR_NO_IR: nothing loaded, outs untouched.
 * HB_ERR_BUSY: set/resize in progress on another thread, block skipped (MonoConvolve.cpp:181-183). */
int hb_matrix_process(hb_matrix *m, const void *const *ins, void *const *outs, uintptr_t num_samples, int accumulate);
int hb_matrix_process_dev(hb_matrix *m, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld,
                          uintptr_t num_samples, int accumulate, void *stream);
/* One matrix on several GPUs of this process -- the N x M Convolver of Convolver.h:25-50 behind ONE handle and ONE host-pointer
 * process call, its input channels dealt to `n_devices` GPUs (device ordinals in `devices`; ins and outs multiples of n_devices),
 * or, for parallel banks (groups > 1), its banks dealt to the GPUs (groups a multiple of n_devices).  Every hb_matrix_* entry point
 * works on the returned handle as on a single-device matrix: set / resize go to the device that holds the pair, process(ins, outs)
 * takes all input rows and returns all output rows.  Device d holds inputs [d * ins / n, (d + 1) * ins / n) against all outputs and
 * sums outputs [d * outs / n, ...): the sum over inputs of NToMonoConvolve.cpp:39-42 crosses devices
 *   - inside the inverse-FFT kernels (peer stores over NVLink, the fused exchange of hb_conv_shard_*) when the scheme is one uniform
 *     FFT size without a head -- calls are pipelined behind the API's own latency of fft_size / 2 as on one device;
 *   - by an owner-side kernel that reads the other devices' partial blocks out of their memory (any scheme, any call size).
 * One host worker thread per device runs the per-device share of every call.  Peer access between all devices is required
 * (HB_ERR_UNSUPPORTED otherwise).  hb_matrix_process_dev is not available on the front; its shards take device pointers. */
int hb_matrix_create_multi(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                           int zero_latency, uint32_t A, uint32_t B, uint32_t C, uint32_t D, const int *devices, uint32_t n_devices);
int hb_matrix_create_latency_multi(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                                   int latency_mode, const int *devices, uint32_t n_devices);
/* the per-device matrices behind a multi-device front (1 and the handle itself on a single-device matrix; borrowed) */
uint32_t hb_matrix_shards(const hb_matrix *m);
hb_matrix *hb_matrix_shard(hb_matrix *m, uint32_t index);
/* how the sum over inputs crosses devices: 0 nothing crosses (one device, or parallel banks), 1 owner-side peer reads, 2 fused into
 * the inverse-FFT kernels */
int hb_matrix_exchange(const hb_matrix *m);
/* introspection: the uniform engines behind the scheme (borrowed handles; the tail is the last) */
uint32_t hb_matrix_parts(const hb_matrix *m);
hb_conv *hb_matrix_part(hb_matrix *m, uint32_t index);
uint32_t hb_matrix_head_taps(const hb_matrix *m);

/* ---------------------------------------------------------------------------------------------
 * One-shot FFT convolution / correlation of two buffers -- replaces spectral_processor<T>::convolve(T *output,
 * in_ptr in1, in_ptr in2, EdgeMode mode) (SpectralProcessor.hpp:169-172, 616-674) with its per-bin
 * product ir_convolve_real (SpectralFunctions.hpp:420-424, 63-84, 274-281), and its neighbours correlate and the
 * complex-input overloads (below).
 * mode: 0 Linear (n1+n2-1 samples), 1 Wrap, 2 WrapCentre, 3 Fold, 4 FoldRepeat (max(n1,n2) samples;
 * SpectralProcessor.hpp:22, 445-481).  Host pointers of the handle's dtype.
 * ------------------------------------------------------------------------------------------- */
typedef struct hb_spectral hb_spectral;

/* spectral_processor(max_fft_size = 32768): SpectralProcessor.hpp:35-42 */
int hb_spectral_create(hb_spectral **out, int dtype, uintptr_t max_fft_size, int device);
void hb_spectral_destroy(hb_spectral *s);
/* set_max_fft_size / max_fft_size: SpectralProcessor.hpp:96-110 */
int hb_spectral_set_max_fft_size(hb_spectral *s, uintptr_t max_fft_size);
uintptr_t hb_spectral_max_fft_size(const hb_spectral *s);
/* convolved_size: SpectralProcessor.hpp:210-213, 546-557 (0 = the operation would not run) */
uintptr_t hb_spectral_convolved_size(const hb_spectral *s, uintptr_t n1, uintptr_t n2, int mode);
/* convolve: *written = samples stored in output (0 when an input is empty or the FFT would exceed the
 * maximum -- the reference silently returns, SpectralProcessor.hpp:651-652) */
int hb_spectral_convolve(hb_spectral *s, void *output, const void *in1, uintptr_t n1, const void *in2, uintptr_t n2,
                         int mode, uintptr_t *written);
/* correlate(T *output, in_ptr in1, in_ptr in2, EdgeMode): SpectralProcessor.hpp:181-184, 483-538 (arrange_correlate),
 * SpectralFunctions.hpp:265-272, 432-436.  Same sizes as convolve (correlated_size = convolved_size, :215-218). */
int hb_spectral_correlate(hb_spectral *s, void *output, const void *in1, uintptr_t n1, const void *in2, uintptr_t n2,
                          int mode, uintptr_t *written);
/* complex inputs -- convolve / correlate(T *r_out, T *i_out, in_ptr r_in1, in_ptr i_in1, in_ptr r_in2, in_ptr i_in2,
 * EdgeMode): SpectralProcessor.hpp:164-167, 176-179, 559-614.  Every plane has its own length (0 = absent plane); the
 * operand length is the longer of its planes.  The reference's Split overload of wrap() takes an offset where its
 * callers pass an end (:421-427 against :437-443), so its Wrap / WrapCentre modes add samples read at or past the end
 * of the transform (unrelated temporary memory); reads past the transform contribute 0 here, everything else follows
 * the reference index for index. */
int hb_spectral_convolve_complex(hb_spectral *s, void *r_out, void *i_out, const void *r_in1, uintptr_t nr1, const void *i_in1, uintptr_t ni1,
                                 const void *r_in2, uintptr_t nr2, const void *i_in2, uintptr_t ni2, int mode, uintptr_t *written);
int hb_spectral_correlate_complex(hb_spectral *s, void *r_out, void *i_out, const void *r_in1, uintptr_t nr1, const void *i_in1, uintptr_t ni1,
                                  const void *r_in2, uintptr_t nr2, const void *i_in2, uintptr_t ni2, int mode, uintptr_t *written);
/* change_phase(T *output, const T *input, uintptr_t size, double phase, double time_multiplier = 1.0): SpectralProcessor.hpp:186-208
 * with ir_phase and minimum_phase_components (SpectralFunctions.hpp:283-336, 405-418): phase 0 = minimum, 0.5 = linear,
 * 1 = maximum phase version of the input.  output receives the FFT size (next power of two of round(size *
 * time_multiplier)) samples; *written = that count.  Unlike the reference (which indexes past its setup) an FFT above the
 * processor's maximum returns HB_ERR_BAD_ARG. */
int hb_spectral_change_phase(hb_spectral *s, void *output, const void *input, uintptr_t size, double phase, double time_multiplier, uintptr_t *written);

/* ---------------------------------------------------------------------------------------------
 * Impulse responses from WAV / AIFF / AIFC files -- replaces the reading half of the reference's AudioFile component
 * (AudioFile/IAudioFile.h:30-54: open, the BaseAudioFile getters, seek, readInterleaved, readChannel), the step before
 * Convolver::set in real use.  Headers are parsed on the host; PCM decoding (8 / 16 / 24 / 32-bit integers of either
 * byte order, 32 / 64-bit floats; de-interleaving) runs on the GPU.
 * ------------------------------------------------------------------------------------------- */
typedef struct hb_audio_info
{
    int32_t file_type;          /* BaseAudioFile::FileType: 0 none, 1 AIFF, 2 AIFC, 3 WAVE (BaseAudioFile.h:19-25) */
    int32_t pcm_format;         /* BaseAudioFile::PCMFormat: 0 int8, 1 int16, 2 int24, 3 int32, 4 float32, 5 float64 (:27-35) */
    int32_t header_big_endian;  /* getHeaderEndianness() == kAudioFileBigEndian */
    int32_t audio_big_endian;   /* getAudioEndianness() == kAudioFileBigEndian */
    uint32_t channels;          /* getChannels() */
    uint32_t frames;            /* getFrames() */
    double sampling_rate;       /* getSamplingRate() */
    uint64_t pcm_offset;        /* byte offset of the first frame (getPCMOffset()) */
    int32_t error_flags;        /* BaseAudioFile::Error bits (:49-62); 0 = readable */
    int32_t reserved;
} hb_audio_info;

/* IAudioFile::open + parseHeader (IAudioFile.cpp:36-53, 375-609).  Returns HB_OK whenever the probe itself ran; what the
 * reference would report through getErrorFlags() is in info->error_flags (e.g. 4 = could not open). */
int hb_audio_probe(const char *path, hb_audio_info *info);
/* seek(first_frame) then readChannel(out, frames, channel) (channel >= 0) or readInterleaved(out, frames) (channel < 0):
 * IAudioFile.cpp:74-115, 613-689.  out: host array of out_dtype (HB_F32 / HB_F64), frames (x channels) elements. */
int hb_audio_read(const char *path, uint32_t first_frame, uint32_t frames, int32_t channel, void *out, int out_dtype, int device);
/* the decode step alone on device memory: d_raw holds `frames` raw interleaved frames as stored in the file; channel >= 0
 * writes that channel to d_out[0 .. frames), channel < 0 writes every channel c to row d_out + c * ld (planar). */
int hb_audio_decode_dev(const hb_audio_info *info, const void *d_raw, uint64_t frames, int32_t channel, void *d_out, uint64_t ld,
                        int out_dtype, int device, void *stream);
/* seek(first_frame) then readRaw(out, frames): IAudioFile.cpp:74-88 -- the frames as the file stores them (host memory,
 * frames x channels x bytes per sample).  Host code only. */
int hb_audio_read_raw(const char *path, uint32_t first_frame, uint32_t frames, void *out);

/* ---- OAudioFile (AudioFile/OAudioFile.h:18-30): the writer.  Host code only; files are byte-identical to the reference's. ----
 * open(path, type, format, channels, sr[, endianness]): OAudioFile.cpp:41-71.  file_type 1 AIFF (written as AIFC, :58), 2 AIFC,
 * 3 WAVE; pcm_format 0 int8 .. 3 int32, 4 float32, 5 float64 (BaseAudioFile.h:23-33); big_endian -1 = the type's default (WAVE
 * little, AIFC big), 0 little (RIFF, or AIFC with little-endian samples -- tagged "NONE" as the reference tags it), 1 big (RIFX).
 * Always returns a handle; a file that could not be opened shows in hb_audio_writer_info (is_open 0, error flag 4). */
typedef struct hb_audio_writer hb_audio_writer;
int hb_audio_writer_open(hb_audio_writer **w, const char *path, int file_type, int pcm_format, uint32_t channels, double rate, int big_endian);
/* writeInterleaved (channel < 0: frames x channels samples) / writeChannel (channel >= 0: frames samples into that channel, the
 * other channels keep what they hold, silence past the old end): OAudioFile.cpp:102-120, 587-682.  in_dtype HB_F32 / HB_F64.
 * Integer formats round half away from zero and wrap instead of clipping (:566-576); 8-bit WAVE is unsigned and clipped (:578-585). */
int hb_audio_writer_write(hb_audio_writer *w, const void *in, int in_dtype, uint32_t frames, int32_t channel);
/* writeRaw: frames as the file stores them (OAudioFile.h:30) */
int hb_audio_writer_write_raw(hb_audio_writer *w, const void *raw, uint32_t frames);
/* seek / getPosition in frames (OAudioFile.cpp:84-96) */
int hb_audio_writer_seek(hb_audio_writer *w, uint32_t frame);
uint32_t hb_audio_writer_position(hb_audio_writer *w);
/* the BaseAudioFile getters of the open file (frames written so far, error flags) and isOpen() */
int hb_audio_writer_info(const hb_audio_writer *w, hb_audio_info *info, int *is_open);
/* close() and release the handle */
void hb_audio_writer_close(hb_audio_writer *w);

/* file -> spectra without a host float array: channel `channel` of the file becomes the impulse response of pair
 * (group, in, out) of a uniform engine (hb_conv_set_ir_dev on the decoded device row). */
int hb_conv_set_ir_file(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const char *path, uint32_t channel, int device);
/* element type of an engine (HB_F32 / HB_F64) */
int hb_conv_dtype(const hb_conv *c);

#ifdef __cplusplus
}
#endif
#endif /* HISSTOOLS_B200_H */
