// tests/cpp/dropin_test.cpp -- a caller written against the REFERENCE's C++ API
// (HISSTools_FFT.h, PartitionedConvolve.h, MonoConvolve.h, NToMonoConvolve.h, Convolver.h), compiled
// against this repo's include/ instead and linked to libhisstools_b200.so.  It writes every result to a
// raw float32 file; tests/test_gpu_cpp_dropin.py regenerates the same inputs and checks them against
// the oracle.  Usage: dropin_test <output file> [audio file] [number of GPUs for the multi-device section]
#include "HISSTools_FFT/HISSTools_FFT.h"
#include "HIRT_Multichannel_Convolution/Convolver.h"
#include "SpectralProcessor.hpp"
#include "AudioFile/IAudioFile.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

static uint32_t lcg_state;
static float lcg() { lcg_state = lcg_state * 1664525u + 1013904223u; return float(lcg_state >> 8) * (1.0f / 8388608.0f) - 1.0f; }
static std::vector<float> noise(size_t n, uint32_t seed) { lcg_state = seed; std::vector<float> v(n); for (auto &x : v) x = lcg(); return v; }
static std::vector<float> decaying(size_t n, uint32_t seed)
{
    std::vector<float> v = noise(n, seed);
    for (size_t k = 0; k < n; k++) v[k] *= 1.0f - float(k) / float(n);
    return v;
}

static FILE *out_file;
static void dump(const std::vector<float> &v) { fwrite(v.data(), sizeof(float), v.size(), out_file); }
static void dump_code(int code) { std::vector<float> v(1, float(code)); dump(v); }

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    out_file = fopen(argv[1], "wb");
    if (!out_file) return 2;

    // 1. FFT family: real forward + inverse of 1024 points, complex forward of 256 points
    {
        FFT_SETUP_F setup;
        hisstools_create_setup(&setup, 10);
        std::vector<float> x = noise(1024, 1), re(512), im(512), y(1024);
        FFT_SPLIT_COMPLEX_F split(re.data(), im.data());
        hisstools_rfft(setup, x.data(), &split, 1000, 10);          // zero-padded from 1000 samples
        dump(re); dump(im);
        hisstools_rifft(setup, &split, y.data(), 10);
        dump(y);
        std::vector<float> cr = noise(256, 2), ci = noise(256, 3);
        FFT_SPLIT_COMPLEX_F c(cr.data(), ci.data());
        hisstools_fft(setup, &c, 8);
        dump(cr); dump(ci);
        hisstools_destroy_setup(setup);

        FFT_SETUP_D setup_d;
        hisstools_create_setup(&setup_d, 8);
        std::vector<double> dr(128), di(128);
        FFT_SPLIT_COMPLEX_D sd(dr.data(), di.data());
        std::vector<float> xf = noise(256, 4);
        hisstools_rfft(setup_d, xf.data(), &sd, 256, 8);             // float input, double setup (HISSTools_FFT.h:208)
        std::vector<float> dre(128), dim(128);
        for (int k = 0; k < 128; k++) { dre[k] = float(dr[k]); dim[k] = float(di[k]); }
        dump(dre); dump(dim);
        hisstools_destroy_setup(setup_d);
    }

    // 2. PartitionedConvolve(512, 3000, 0, 0): ragged call sizes, error codes
    {
        HISSTools::PartitionedConvolve pc(512, 3000, 0, 0);
        std::vector<float> x = noise(6000, 10), y(6000, 7.f), ir = decaying(3000, 11);
        dump_code(pc.process(x.data(), y.data(), 100) ? 1 : 0);     // no IR yet: false, y untouched
        dump_code(int(y[0]));
        pc.setResetOffset(0);
        dump_code(pc.set(ir.data(), ir.size()));
        dump_code(pc.setFFTSize(16));                                 // CONVOLVE_ERR_FFT_SIZE_OUT_OF_RANGE
        const size_t sizes[] = {1, 255, 256, 257, 1000, 31};
        size_t pos = 0, k = 0;
        while (pos < x.size())
        {
            size_t n = std::min(sizes[k++ % 6], x.size() - pos);
            pc.process(x.data() + pos, y.data() + pos, n);
            pos += n;
        }
        dump(y);
    }

    // 3. MonoConvolve(5000, kLatencyShort) and a custom uniform scheme with accumulate
    {
        HISSTools::MonoConvolve mc(5000, kLatencyShort);
        std::vector<float> x = noise(4096, 20), y(4096), t(4096), ir = decaying(5000, 21);
        mc.setResetOffset(0);
        dump_code(mc.set(ir.data(), ir.size(), false));
        for (size_t pos = 0; pos < x.size(); pos += 128) mc.process(x.data() + pos, t.data(), y.data() + pos, 128);
        dump(y);
        HISSTools::MonoConvolve moved(std::move(mc));
        bool threw = false;
        try { HISSTools::MonoConvolve bad(1000, false, 1024, 256); } catch (std::runtime_error &) { threw = true; }
        dump_code(threw ? 1 : 0);
    }

    // 4. NToMonoConvolve(3, 2000, kLatencyMedium)
    {
        HISSTools::NToMonoConvolve nm(3, 2000, kLatencyMedium);
        std::vector<std::vector<float>> xs, irs;
        for (uint32_t i = 0; i < 3; i++) { xs.push_back(noise(4096, 30 + i)); irs.push_back(decaying(2000, 40 + i)); }
        for (uint32_t i = 0; i < 3; i++) dump_code(nm.set(i, irs[i].data(), 2000, false));
        dump_code(nm.set(3, irs[0].data(), 2000, false));              // CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        std::vector<float> y(4096), t(512);
        for (size_t pos = 0; pos < 4096; pos += 512)
        {
            const float *ins[3] = { xs[0].data() + pos, xs[1].data() + pos, xs[2].data() + pos };
            nm.process(ins, y.data() + pos, t.data(), 512, 3);
        }
        dump(y);
    }

    // 5. Convolver(3, 2, kLatencyZero), 20000-tap IRs through resize, float and double I/O
    {
        HISSTools::Convolver cv(3, 2, kLatencyZero);
        std::vector<std::vector<float>> xs, irs;
        for (uint32_t i = 0; i < 3; i++) xs.push_back(noise(8192, 50 + i));
        for (uint32_t p = 0; p < 6; p++) irs.push_back(decaying(20000, 60 + p));
        dump_code(cv.set(0, 0, irs[0].data(), 20000, false));          // default room is 16384 taps: CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL
        for (uint32_t o = 0; o < 2; o++)
            for (uint32_t i = 0; i < 3; i++) dump_code(cv.set(i, o, irs[o * 3 + i].data(), 20000, true));
        dump_code(cv.set(0, 2, irs[0].data(), 20000, true));           // CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE
        std::vector<float> y0(8192), y1(8192);
        for (size_t pos = 0; pos < 8192; pos += 64)
        {
            const float *ins[3] = { xs[0].data() + pos, xs[1].data() + pos, xs[2].data() + pos };
            float *outs[2] = { y0.data() + pos, y1.data() + pos };
            cv.process(ins, outs, 3, 2, 64);
        }
        dump(y0); dump(y1);
        // double I/O on a fresh stream
        cv.reset();
        std::vector<std::vector<double>> xd(3, std::vector<double>(1024));
        for (uint32_t i = 0; i < 3; i++) for (size_t k = 0; k < 1024; k++) xd[i][k] = xs[i][k];
        std::vector<double> d0(1024), d1(1024);
        const double *dins[3] = { xd[0].data(), xd[1].data(), xd[2].data() };
        double *douts[2] = { d0.data(), d1.data() };
        cv.process(dins, douts, 3, 2, 1024);
        std::vector<float> f0(d0.begin(), d0.end());
        dump(f0);
    }

    // 6. Convolver(4, kLatencyShort): parallel channels
    {
        HISSTools::Convolver cv(4, kLatencyShort);
        std::vector<std::vector<float>> xs, irs, ys(4, std::vector<float>(2048));
        for (uint32_t i = 0; i < 4; i++) { xs.push_back(noise(2048, 70 + i)); irs.push_back(decaying(1500, 80 + i)); }
        for (uint32_t i = 0; i < 4; i++) dump_code(cv.set(i, i, irs[i].data(), 1500, false));
        dump_code(cv.set(1, 0, irs[0].data(), 1500, false));           // parallel mode: in must equal out
        const float *ins[4] = { xs[0].data(), xs[1].data(), xs[2].data(), xs[3].data() };
        float *outs[4] = { ys[0].data(), ys[1].data(), ys[2].data(), ys[3].data() };
        cv.process(ins, outs, 4, 4, 2048);
        for (auto &y : ys) dump(y);
    }

    // 7. spectral_processor<float>::convolve, Linear and WrapCentre
    {
        spectral_processor<float> sp;
        std::vector<float> a = noise(900, 90), b = decaying(250, 91), lin(1149), wc(900);
        dump_code(int(sp.convolved_size(900, 250, spectral_processor<float>::EdgeMode::Linear)));
        sp.convolve(lin.data(), spectral_processor<float>::in_ptr(a.data(), a.size()), spectral_processor<float>::in_ptr(b.data(), b.size()),
                    spectral_processor<float>::EdgeMode::Linear);
        dump(lin);
        sp.convolve(wc.data(), spectral_processor<float>::in_ptr(a.data(), a.size()), spectral_processor<float>::in_ptr(b.data(), b.size()),
                    spectral_processor<float>::EdgeMode::WrapCentre);
        dump(wc);
    }

    // 8. spectral_processor<float>: correlate (Wrap), complex-input convolve (Linear), change_phase (minimum phase)
    {
        typedef spectral_processor<float> SP;
        SP sp;
        std::vector<float> a = noise(700, 92), b = decaying(300, 93), ai = noise(350, 94), bi = noise(300, 95);
        std::vector<float> corr(700), cr(999), ci(999), mp(1024);
        dump_code(int(sp.correlated_size(700, 300, SP::EdgeMode::Wrap)));
        sp.correlate(corr.data(), SP::in_ptr(a.data(), a.size()), SP::in_ptr(b.data(), b.size()), SP::EdgeMode::Wrap);
        dump(corr);
        sp.convolve(cr.data(), ci.data(), SP::in_ptr(a.data(), a.size()), SP::in_ptr(ai.data(), ai.size()), SP::in_ptr(b.data(), b.size()),
                    SP::in_ptr(bi.data(), bi.size()), SP::EdgeMode::Linear);
        dump(cr); dump(ci);
        sp.change_phase(mp.data(), b.data(), 300, 0.0, 3.0);       // FFT 1024
        dump(mp);
    }

    // 9. HISSTools::IAudioFile: header getters, readChannel after a seek, readInterleaved
    if (argc > 2)
    {
        HISSTools::IAudioFile f(argv[2]);
        dump_code(f.isOpen() ? 1 : 0);
        dump_code(f.getErrorFlags());
        dump_code(int(f.getFileType())); dump_code(int(f.getPCMFormat())); dump_code(f.getChannels()); dump_code(int(f.getFrames()));
        dump_code(int(f.getSamplingRate())); dump_code(f.getBitDepth()); dump_code(int(f.getFrameByteCount()));
        std::vector<float> ch(100), inter(size_t(f.getFrames()) * f.getChannels());
        f.seek(17);
        f.readChannel(ch.data(), 100, 1);
        dump_code(int(f.getPosition()));
        dump(ch);
        f.seek();
        f.readInterleaved(inter.data(), f.getFrames());
        dump(inter);
        HISSTools::IAudioFile missing("/nonexistent/file.wav");
        dump_code(missing.isOpen() ? 1 : 0);
        dump_code(missing.getErrorFlags());
    }

    // 10. (argv[3] = number of GPUs, >= 2) the same Convolver class dealt to several GPUs of this process: a uniform 8 x 8
    // matrix (exchange fused into the inverse-FFT kernels), Convolver(4, 4, kLatencyShort) (owner-side peer reads) and four
    // parallel channels, all through the reference's process(ins, outs, ...) with host pointers and ragged call sizes
    const int ndev = argc > 3 ? atoi(argv[3]) : 0;
    if (ndev >= 2)
    {
        std::vector<int> devices;
        for (int d = 0; d < ndev; d++) devices.push_back(d);
        {
            const uint32_t N = 8;
            HISSTools::Convolver cv(N, N, 3000, devices, false, 512);
            cv.setResetOffset(0);
            std::vector<std::vector<float>> xs, ys(N, std::vector<float>(6000));
            for (uint32_t i = 0; i < N; i++) xs.push_back(noise(6000, 100 + i));
            for (uint32_t o = 0; o < N; o++)
                for (uint32_t i = 0; i < N; i++) { std::vector<float> ir = decaying(3000, 200 + o * N + i); dump_code(cv.set(i, o, ir.data(), 3000, false)); }
            const size_t sizes[] = {256, 256, 100, 1024, 1, 411, 256};
            size_t pos = 0, k = 0;
            while (pos < 6000)
            {
                const size_t n = std::min(sizes[k++ % 7], 6000 - pos);
                const float *ins[N]; float *outs[N];
                for (uint32_t i = 0; i < N; i++) { ins[i] = xs[i].data() + pos; outs[i] = ys[i].data() + pos; }
                cv.process(ins, outs, N, N, n);
                pos += n;
            }
            dump_code(hb_matrix_exchange(cv.handle()));
            for (auto &y : ys) dump(y);
        }
        const std::vector<int> dev4(devices.begin(), devices.begin() + std::min(ndev, 4));      // four channels: at most four devices
        {
            const uint32_t N = 4;
            HISSTools::Convolver cv(N, N, kLatencyShort, dev4);
            cv.setResetOffset(0);
            std::vector<std::vector<float>> xs, ys(N, std::vector<float>(4096));
            for (uint32_t i = 0; i < N; i++) xs.push_back(noise(4096, 300 + i));
            for (uint32_t o = 0; o < N; o++)
                for (uint32_t i = 0; i < N; i++) { std::vector<float> ir = decaying(9000, 400 + o * N + i); dump_code(cv.set(i, o, ir.data(), 9000, false)); }
            for (size_t pos = 0; pos < 4096; pos += 512)
            {
                const float *ins[N]; float *outs[N];
                for (uint32_t i = 0; i < N; i++) { ins[i] = xs[i].data() + pos; outs[i] = ys[i].data() + pos; }
                cv.process(ins, outs, N, N, 512);
            }
            dump_code(hb_matrix_exchange(cv.handle()));
            for (auto &y : ys) dump(y);
        }
        {
            const uint32_t N = 4;
            HISSTools::Convolver cv(N, kLatencyMedium, dev4);
            cv.setResetOffset(0);
            std::vector<std::vector<float>> xs, ys(N, std::vector<float>(4096));
            for (uint32_t i = 0; i < N; i++) { xs.push_back(noise(4096, 500 + i)); std::vector<float> ir = decaying(3000, 600 + i); dump_code(cv.set(i, i, ir.data(), 3000, false)); }
            const float *ins[N]; float *outs[N];
            for (uint32_t i = 0; i < N; i++) { ins[i] = xs[i].data(); outs[i] = ys[i].data(); }
            cv.process(ins, outs, N, N, 4096);
            dump_code(hb_matrix_exchange(cv.handle()));
            for (auto &y : ys) dump(y);
        }
    }

    fclose(out_file);
    printf("dropin_test: %llu kernel launches\n", (unsigned long long) hb_launch_count());
    return 0;
}
