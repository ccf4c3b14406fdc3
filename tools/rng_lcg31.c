// tools/rng_lcg31.c -- offline research tool: exhaustive search of ALL states of the minstd engines (state < 2^31 - 1) for the
// README-GIF fork stream.  gcc -O2 -fopenmp -o /tmp/lcg31 tools/rng_lcg31.c && /tmp/lcg31
// draw rule of uniform_int_distribution(0,1): bit = x / ((max - max%2)/2), or the low bit; plain and complemented.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static const char *TAPE = "01111110010111111111011001111101011001000110100110011001010000111010100111100010110000000011000111001000101011010101101101110111010000101101101001001000010011010001100101110011100110110110110100000101011010110011000100001000110001101101001110000011101110001111000100101101100011001000011110101001001100111100111110010100111010010101111101001010111000001101101001111110111101011100100000111011000001";
int main() {
  const uint64_t mults[2] = {16807, 48271};
  const uint64_t M = 2147483647ull;
  int n = strlen(TAPE);
  for (int mi = 0; mi < 2; mi++) {
    uint64_t a = mults[mi];
    long found = 0;
    #pragma omp parallel for schedule(dynamic, 1<<16) reduction(+:found)
    for (uint64_t s0 = 1; s0 < M; s0++) {
      for (int variant = 0; variant < 4; variant++) {  // bit rule: top-ish / low bit; complement
        uint64_t s = s0; int t = 0;
        for (; t < n; t++) {
          uint64_t x;
          do { s = s * a % M; x = s; } while (x >= 2147483646ull);
          int b = (variant & 1) ? (int)(x & 1) : (int)(x / 1073741823ull);
          if (variant & 2) b ^= 1;
          if (b != TAPE[t] - '0') break;
        }
        if (t >= 60) { found++; printf("mult %lu state %lu variant %d matched %d bits\n", a, s0, variant, t); }
      }
    }
    printf("mult %lu: %ld candidates with >= 60 bits\n", a, found);
  }
  return 0;
}
