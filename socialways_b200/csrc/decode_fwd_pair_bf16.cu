// The bf16 build of the CTA-pair decode kernel: the same schedule, tile prologue and epilogues as decode_fwd_pair.cu on single bf16
// operands (one tcgen05.mma per product instead of three; the x-feedback block keeps its hi + lo parts).  Entry point
// sw_decode_fwd_pair_bf16, pack from packing.pack_decoder_pair(..., bf16=True).  The fast mode BASELINE configs[2] ("bf16") names.
#define SW_PAIR_BF16 1
#include "decode_fwd_pair.cu"
