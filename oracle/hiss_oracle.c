/* oracle/hiss_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the HISSTools_Library partitioned-convolution path (the FFT
 * conventions of HISSTools_FFT, PartitionedConvolve, TimeDomainConvolve, MonoConvolve's
 * partition scheme and spectral_processor::convolve), in float and double.  It is the CHECKER
 * for the CUDA product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library never links or calls it.
 *
 * Parity of this oracle is PINNED: tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference compiled in place (oracle/_ref, built by oracle/Makefile) and
 * tests/golden/ holds fixtures generated from that reference (tests/golden/make_golden.py) so the
 * pin also holds on machines without /root/reference.  The reference's own tests hold only one
 * known-answer check on this path (zip/unzip exactness, "- Test/FFT_Tester/FFT_Tester/main.cpp":201-250),
 * which tests/test_oracle_golden.py reproduces.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#include "hiss_oracle.h"

#define T float
#define SUF _f32
#include "hiss_oracle_impl.inc"
#undef T
#undef SUF

#define T double
#define SUF _f64
#include "hiss_oracle_impl.inc"
#undef T
#undef SUF
