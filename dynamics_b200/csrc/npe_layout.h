// Shared host/device constants: model dimensions, the flat (reference) parameter layout and the padded
// shared-memory weight layout used by the sm_100a kernels.
//
// Flat layout == nn.Module:getParameters() of init_network (reference src/lua/npe.lua:14-61,
// modules.lua:10-26,46-88) with -nbrhd; see include/npe_cuda.h.
#pragma once

namespace npe {

constexpr int D = 22;        // object_dim                (config.lua:25)
constexpr int IN = 44;       // input_dim = 2 * 22        (main.lua:122)
constexpr int H = 50;        // rnn_dim                   (main.lua:37)
constexpr int H2 = 25;       // encoder half width        (modules.lua:17,19)
constexpr int NL = 5;        // layers                    (main.lua:40)
constexpr int OUT = 22;      // out_dim                   (main.lua:123)
constexpr int DEC_IN = 94;   // rnn_dim + input_dim       (modules.lua:54)

// ---- flat offsets (floats) ----
constexpr int OFF_ENC_T = 0;
constexpr int OFF_ENC_C = 1100;
constexpr int OFF_PAIR = 2200;      // + 2500 * l
constexpr int OFF_DEC1_W = 14700;   // 50 x 94
constexpr int OFF_DEC1_B = 19400;
constexpr int OFF_DEC2_W = 19450;   // dec layers 2..4: W at 19450 + 2550*(i-2), b at W + 2500
constexpr int OFF_DEC5_W = 27100;   // 22 x 50
constexpr int OFF_DEC5_B = 28200;
constexpr int NPARAM = 28222;
constexpr int NPARAM_PAIR = 14700;  // [0, 14700) = encoder + pairwise layers; [14700, 28222) = decoder
__host__ __device__ constexpr int off_dec_w(int i) { return OFF_DEC2_W + 2550 * (i - 2); }  // i = 2..4
__host__ __device__ constexpr int off_dec_b(int i) { return off_dec_w(i) + 2500; }

// ---- padded shared-memory layout ----
// Activations are row-major [row][LD]; weights keep the reference (out,in) orientation [n][LDW] so that ONE copy
// serves the forward (inner-product form, k contiguous) and the input-gradient pass (outer-product form, the
// output index contiguous).  LD = 56 (K <= 52) / 104 (decoder layer 1, K = 96): LD mod 32 in {8, 24} makes the
// float4 accesses of 4 consecutive rows x 2 k-chunks hit 8 distinct 16-byte bank groups.
// Decoder biases are folded in as column 50 of the weight rows against a constant-1 activation column 50
// (row 50 of each hidden decoder layer is the unit vector e_50, which regenerates the constant).
constexpr int LD = 56;
constexpr int LD1 = 104;
constexpr int ZCOL_F = 52;          // decoder input z = [s(0..49) | 1 | 0 | this_past(52..95)]
constexpr int ONE_COL = 50;

// Weight tiles are padded to the row counts the thread mapping touches (output column n = cg + 8 i, cg < 8):
// hidden layers 56 rows, encoders 32 rows, decoder output layer 24 rows; pad rows/columns are zero.
constexpr int SW_ENC_T = 0;                       // 32 x 56
constexpr int SW_ENC_C = SW_ENC_T + 32 * LD;      // 32 x 56
constexpr int SW_PAIR = SW_ENC_C + 32 * LD;       // 5 x (56 x 56)
constexpr int SW_PAIR_SZ = 56 * LD;
constexpr int SW_PAIR_TOTAL = SW_PAIR + NL * SW_PAIR_SZ;          // 19264 floats = 77056 B

constexpr int SW_DEC1 = 0;                        // 56 x 104
constexpr int SW_DEC2 = SW_DEC1 + 56 * LD1;       // 3 x (56 x 56)
constexpr int SW_DEC5 = SW_DEC2 + 3 * SW_PAIR_SZ; // 24 x 56
constexpr int SW_DEC_TOTAL = SW_DEC5 + 24 * LD;   // 16576 floats = 66304 B

// ---- per-CTA gradient partials: TILE-MAJOR layout ----
// A CTA writes its weight-gradient register tiles (4 x 4 per thread) as they are: 16 consecutive floats per thread per
// accumulator array, i.e. four 16-byte stores, coalesced over the warp, no index arithmetic on the critical path.  The
// reduce kernels (one parameter per thread) translate a flat parameter index into this layout (partial_index).
constexpr int PT_PAIR = 0;                          // 5 layers x 169 tiles (13 x 13) x 16
constexpr int PT_LAYER = 169 * 16;
constexpr int PT_ENC_C = PT_PAIR + NL * PT_LAYER;   // 77 tiles (7 x 11) x 16
constexpr int PT_ENC_T = PT_ENC_C + 77 * 16;
constexpr int PT_DEC1A = PT_ENC_T + 77 * 16;        // decoder layer 1: tiles 0..255 of 13 x 24
constexpr int PT_DEC1B = PT_DEC1A + 256 * 16;       //                  tiles 256..311
constexpr int PT_DEC2 = PT_DEC1B + 56 * 16;         // decoder layers 2..4: 169 tiles each
constexpr int PT_DEC5 = PT_DEC2 + 3 * PT_LAYER;     // decoder layer 5: 78 tiles (6 x 13)
constexpr int PT_TOTAL = PT_DEC5 + 78 * 16;         // 30336 floats per CTA

constexpr int TILE = 64;            // rows (foci or pairs) per shared-memory tile
constexpr int NT = 256;             // threads per CTA
constexpr int MAXOBJ = 64;          // objects per scene

}  // namespace npe
