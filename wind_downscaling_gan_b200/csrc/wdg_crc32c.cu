// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) for the TensorFlow checkpoint-V2 writer / reader
// (tf_checkpoint.py): the SSTable block trailers of `<prefix>.index` and the per-tensor checksums of BundleEntryProto
// are `mask(crc32c(bytes))`.  Host code, slicing-by-8 tables built on first use.
#include <stddef.h>
#include <stdint.h>

#include "../../include/wdg.h"

namespace {
uint32_t g_tab[8][256];
bool g_init = false;
void init_tables() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xff];
  g_init = true;
}
}  // namespace

extern "C" uint32_t wdg_crc32c(uint32_t crc, const void* data, size_t n) {
  if (!g_init) init_tables();
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xff]; --n; }
  while (n >= 8) {
    const uint64_t v = *(const uint64_t*)p ^ c;      // little-endian host (x86-64 / aarch64)
    c = g_tab[7][v & 0xff] ^ g_tab[6][(v >> 8) & 0xff] ^ g_tab[5][(v >> 16) & 0xff] ^ g_tab[4][(v >> 24) & 0xff] ^
        g_tab[3][(v >> 32) & 0xff] ^ g_tab[2][(v >> 40) & 0xff] ^ g_tab[1][(v >> 48) & 0xff] ^ g_tab[0][(v >> 56) & 0xff];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xff];
  return ~c;
}
