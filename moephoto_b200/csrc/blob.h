// Packed weight blob (host memory, little endian) produced by moephoto_b200/weights.py and consumed by
// moe_model_load.  One header, a directory, then 256-byte aligned sections.
#pragma once
#include <cstdint>

namespace moe {

constexpr uint32_t kBlobMagic = 0x42454F4Du;   // "MOEB"
constexpr uint32_t kBlobVersion = 2;

enum BlobSection : uint32_t {
  SEC_FIRST_W = 1,   // float [9][64]          conv_input.weight, tap-major, output channel minor
  SEC_SCALARS = 2,   // float [32]             [0] relu.weight ; [1+l] trunk layer l: PReLU slope (conv_1)
                     //                        or scale (conv_2), unused for conv_input2 ;
                     //                        [14 + 4*branch + stage] PReLU slope of an upsample block
  SEC_TRUNK_IMG = 3, // index l = 0..12        conv_input2, then conv_1 / conv_2 of ARSB 1..6 (MoeNet_lite2: l = 0..6, LB 1..3):
                     //                        fp16 [9 taps][64 out][64 in] as 128-byte rows with the
                     //                        16-byte chunk index XOR-swizzled by (row & 7)  (73 728 B)
  SEC_UP_IMG = 4,    // index 4*branch+stage   r*r such images; image q=(i,j) holds output channels
                     //                        c*r*r + i*r + j, c = 0..63 (PixelShuffle sub-pixel (i,j))
  SEC_UP_BIAS = 5,   // index 4*branch+stage   float [r*r][64] in the same order
  SEC_HEAD_W = 6,    // index branch (0=u,1=R) float [9][64] tap-major, input channel minor
  SEC_FRM = 7,       // index LB block 0..2    MoeNet_lite2's FRM gate: float w0[3][64], b0[4], w1[64][4], b1[64]
};

struct BlobHeader {
  uint32_t magic, version;
  uint32_t arch;       // MoeArch
  uint32_t feat;       // real filter count (64 or 48); all tensors are zero-padded to 64
  uint32_t n_up;       // upsample blocks per branch: 0 (dn), 1 (x2, x3), 2 (x4), up to 3 (MoeNet_lite2 x8)
  uint32_t r;          // PixelShuffle factor of each block (2 or 3), 0 if n_up == 0
  uint32_t n_sections;
  uint32_t reserved;
};

struct BlobEntry {
  uint32_t kind, index;
  uint64_t offset, nbytes;   // offset from the start of the blob
};

}  // namespace moe
