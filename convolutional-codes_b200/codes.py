"""Code descriptions: the reference's codebook plus the BASELINE codes it does not carry.

Mirror of ``common/codebook.c:14-68`` / ``get_code`` (``:110-120``): each entry is
(symlen_out n, constr_len K, block_len L, polynomials).  Polynomials are left-aligned 64-bit
words, bit 63 = tap on the newest input bit (``common/codebook.c:14-17``).  The reference has
no K=7 / K=9 / K=15 entry (SURVEY.md finding 4); those are supplied here as extra values.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

PARITY_REF = 0      # labels as the reference's 0x6996 expression evaluates (bit-parity with the reference)
PARITY_LINEAR = 1   # true GF(2) parity (textbook code)
SOFT_F32 = 0
HARD_U16 = 1
SOFT_U8 = 2        # quantised soft: small integer branch metrics, uint8 [F][T][M]
FORCE_GENERIC = 1
PREFER_32BIT = 4       # hard decisions: 32-bit register kernel instead of the packed 16-bit one
PREFER_LANEGROUP = 2   # K = 8 .. 12: shuffle-exchange lane-group kernel instead of the register-tile kernel
NO_JIT = 8             # unregistered polynomials: table-driven kernels instead of a kernel compiled at run time (NVRTC)


def left_align(octal: str, K: int) -> int:
    """'171' (octal), K=7 -> 0xF200000000000000"""
    v = int(octal, 8)
    assert v < (1 << K)
    return (v << (64 - K)) & 0xFFFFFFFFFFFFFFFF


@dataclass(frozen=True)
class CodeParam:
    """Same meaning as ``struct code_param`` (common/include/code.h:9-19), widened."""
    symlen_out: int
    constr_len: int
    block_len: int
    polynomials: Tuple[int, ...]
    name: str = field(default="", compare=False)

    @property
    def n(self): return self.symlen_out
    @property
    def K(self): return self.constr_len
    @property
    def L(self): return self.block_len
    @property
    def S(self): return 1 << (self.constr_len - 1)
    @property
    def T(self): return self.block_len + self.constr_len - 1
    @property
    def M(self): return 1 << self.symlen_out
    @property
    def out_bytes(self): return (self.T + 7) // 8
    @property
    def payload_bytes(self): return (self.block_len + 7) // 8

    def with_block_len(self, L: int) -> "CodeParam":
        return CodeParam(self.symlen_out, self.constr_len, int(L), self.polynomials, self.name)


# common/codebook.c — index 4 (K=32 WSPR code) is kept for completeness; Viterbi cannot run it.
_CODEBOOK = (
    CodeParam(2, 3, 40, (0xA000000000000000, 0x6000000000000000), "codebook0 (101),(011)"),
    CodeParam(2, 4, 40, (0xB000000000000000, 0xE000000000000000), "codebook1 (1011),(1110)"),
    CodeParam(2, 5, 40, (0xA800000000000000, 0xF000000000000000), "codebook2 (10101),(11110)"),
    CodeParam(2, 6, 40, (0xB400000000000000, 0xE800000000000000), "codebook3 (101101),(111010)"),
    CodeParam(2, 32, 50, (0x8ACA0B4F00000000, 0xE23C862700000000), "codebook4 WSPR K=32"),
    CodeParam(3, 3, 40, (0xA000000000000000, 0xC000000000000000, 0x2000000000000000), "codebook5 (101),(110),(001)"),
)


def get_code(index: int) -> CodeParam:
    """``get_code`` of the reference (common/codebook.c:110-120)."""
    return _CODEBOOK[index]


# BASELINE.json configs that are not in the reference codebook
K7_R12 = CodeParam(2, 7, 1024, (left_align("171", 7), left_align("133", 7)), "K7 (171,133)")
K7_R13 = CodeParam(3, 7, 1024, (left_align("133", 7), left_align("171", 7), left_align("165", 7)), "K7 (133,171,165)")
K8_R12 = CodeParam(2, 8, 1024, (left_align("371", 8), left_align("247", 8)), "K8 (371,247)")
K9_R12 = CodeParam(2, 9, 1024, (left_align("561", 9), left_align("753", 9)), "K9 (561,753)")
K10_R12 = CodeParam(2, 10, 1024, (left_align("1167", 10), left_align("1545", 10)), "K10 (1167,1545)")
K11_R12 = CodeParam(2, 11, 1024, (left_align("2335", 11), left_align("3661", 11)), "K11 (2335,3661)")
K12_R12 = CodeParam(2, 12, 1024, (left_align("4335", 12), left_align("5723", 12)), "K12 (4335,5723)")
K13_R12 = CodeParam(2, 13, 1024, (left_align("10533", 13), left_align("17661", 13)), "K13 (10533,17661)")
K14_R12 = CodeParam(2, 14, 1024, (left_align("21675", 14), left_align("27123", 14)), "K14 (21675,27123)")
K15_R12 = CodeParam(2, 15, 1024, (left_align("46321", 15), left_align("51271", 15)), "K15 (46321,51271)")

# noise scaling of the AWGN driver: sigma = 2^-1/2 * 10^(-EbN0_dB/20)  (AWGN-channel/main.c:153-161)
def awgn_sigma(ebn0_db: float) -> float:
    return (2.0 ** -0.5) * 10.0 ** (-ebn0_db / 20.0)
