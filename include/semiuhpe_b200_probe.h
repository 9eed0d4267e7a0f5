/* semiuhpe_b200_probe.h -- measurement entry points of libsemiuhpe_b200.so.
 *
 * Not part of the drop-in boundary (include/semiuhpe_b200.h): these launch synthetic instruction
 * streams that bench.py and profiles/ use to measure the FP32-pipe peak the roofline divides by.
 * They replace nothing in the reference.
 */
#ifndef SEMIUHPE_B200_PROBE_H
#define SEMIUHPE_B200_PROBE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FP32-pipe probe used by bench.py for the roofline denominator: launches `blocks` CTAs of
 * 256 threads, each thread running `iters` rounds of 8 independent dependent-FMA chains.
 * variant 0: scalar FFMA, 1: packed fma.rn.f32x2, 2: FFMA + 1 MUFU.EX2 per 8 FMA,
 * 3: packed and scalar chains interleaved 1:1, 4 / 5: every packed FMA followed by one LOP3 / IADD
 * (does a 2-cycle FFMA2 leave an issue slot for the ALU pipe?).
 * 6 / 7: packed FMA with an immediate addend / a broadcast scalar multiplier (K2's Horner operand forms).
 * 8 / 9: operand-bandwidth probes -- every packed FMA reads distinct 64-bit register pairs, none shared with its
 *        neighbours: three pairs (x = y*z + x, the form of K2L's sums) / two pairs and an immediate (K2's Horner step).
 * 10-13: issue-mix probes -- 8 packed chains with, per FFMA2, one FSEL (10) / one FMNMX (11), or per 8 FFMA2 one
 *        LDS.128 (12) / one MUFU.EX2 (13): what an instruction of another pipe costs next to the packed stream.
 * FMAs executed = blocks*256*iters*64*2 for variants >= 4 (the extra instructions are not counted), {1, 2, 1, 3} for 0-3.
 * variant 100+v: K2 pass-body probe, 16 warps per CTA each running `iters` 128-node passes of run type
 * v&3 (bit 2: no table loads, bit 3: no MUFU, bit 4: no slot mask); 46 packed FMA-pipe ops per pass. */
int suhpe_fp32_probe(float* sink, int32_t variant, int32_t iters, int32_t blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMIUHPE_B200_PROBE_H */
