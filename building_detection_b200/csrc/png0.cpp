// png0.cpp -- multi-threaded level-0 ("stored") PNG encoder for 8-bit grey masks (host code, no GPU, no zlib).
//
// The reference hands its masks from stage to stage as PNG files written with cv.imwrite(..., [IMWRITE_PNG_COMPRESSION,
// 0]) (predict.py:115, model_fuse.py:350) and base64-encodes the result file for the POST answer (buildAPI.py:122-123).
// At 20 000 x 20 000 px that is 400 MB per mask: libpng (single thread, CRC + Adler byte by byte) needs seconds, more
// than the GPU needs for the whole scene.  A stored PNG is a fixed skeleton around the raw rows, so the bytes can be
// produced in parallel: every thread copies a range of 65 535-byte deflate blocks and computes the CRC-32 of its part
// of the IDAT payload and the Adler-32 of its raw bytes; the partial checksums are combined like zlib's
// crc32_combine / adler32_combine.  Output is byte-identical to building_detection_b200/png0.py (tests/test_host_io.py).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/bd_b200.h"

namespace {

uint32_t g_crc[8][256];
bool g_crc_ready = false;
void crc_init() {
  if (g_crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    g_crc[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc[t][i] = (g_crc[t - 1][i] >> 8) ^ g_crc[0][g_crc[t - 1][i] & 0xFFu];
  g_crc_ready = true;
}
// CRC-32 (IEEE, reflected), slicing-by-8; crc in / out are the plain register value (pre / post conditioning inside)
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = g_crc[0][(c ^ *p++) & 0xFFu] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    const uint32_t lo = static_cast<uint32_t>(v) ^ c, hi = static_cast<uint32_t>(v >> 32);
    c = g_crc[7][lo & 0xFFu] ^ g_crc[6][(lo >> 8) & 0xFFu] ^ g_crc[5][(lo >> 16) & 0xFFu] ^ g_crc[4][lo >> 24] ^
        g_crc[3][hi & 0xFFu] ^ g_crc[2][(hi >> 8) & 0xFFu] ^ g_crc[1][(hi >> 16) & 0xFFu] ^ g_crc[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return ~c;
}
// crc of A||B from crc(A), crc(B), len(B): multiply crc(A) by x^(8 len(B)) in GF(2)[x] / P (zlib's method)
uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t s = 0;
  while (vec) { if (vec & 1u) s ^= *mat; vec >>= 1; ++mat; }
  return s;
}
void gf2_square(uint32_t* sq, const uint32_t* mat) { for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]); }
uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;
  uint32_t row = 1;
  for (int n = 1; n < 32; ++n) { odd[n] = row; row <<= 1; }
  gf2_square(even, odd);
  gf2_square(odd, even);
  do {
    gf2_square(even, odd);
    if (len2 & 1u) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1u) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}
constexpr uint32_t ADLER_MOD = 65521u;
uint32_t adler32_update(uint32_t adler, const uint8_t* p, size_t n) {
  uint32_t a = adler & 0xFFFFu, b = adler >> 16;
  while (n) {
    const size_t k = std::min<size_t>(n, 5552);
    for (size_t i = 0; i < k; ++i) { a += p[i]; b += a; }
    a %= ADLER_MOD; b %= ADLER_MOD;
    p += k; n -= k;
  }
  return (b << 16) | a;
}
uint32_t adler32_combine(uint32_t ad1, uint32_t ad2, uint64_t len2) {
  const uint32_t rem = static_cast<uint32_t>(len2 % ADLER_MOD);
  uint32_t sum1 = ad1 & 0xFFFFu;
  uint32_t sum2 = static_cast<uint32_t>((static_cast<uint64_t>(rem) * sum1) % ADLER_MOD);
  sum1 += (ad2 & 0xFFFFu) + ADLER_MOD - 1;
  sum2 += (ad1 >> 16) + (ad2 >> 16) + ADLER_MOD - rem;
  if (sum1 >= ADLER_MOD) sum1 -= ADLER_MOD;
  if (sum1 >= ADLER_MOD) sum1 -= ADLER_MOD;
  if (sum2 >= (ADLER_MOD << 1)) sum2 -= (ADLER_MOD << 1);
  if (sum2 >= ADLER_MOD) sum2 -= ADLER_MOD;
  return sum1 | (sum2 << 16);
}
void put32be(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

constexpr uint64_t BLOCK = 65535;

}  // namespace

extern "C" {

// bytes bd_png0_encode writes for an (h, w) mask
size_t bd_png0_size(int h, int w) {
  const uint64_t n = static_cast<uint64_t>(h) * (static_cast<uint64_t>(w) + 1);
  const uint64_t nblk = std::max<uint64_t>(1, (n + BLOCK - 1) / BLOCK);
  return static_cast<size_t>(8 + 25 + 12 + (2 + n + 5 * nblk + 4) + 12);
}

// mask_host: (h, w) u8, rows contiguous.  out_host: at least bd_png0_size(h, w) bytes.  threads <= 0: all cores (max 32).
int bd_png0_encode(const uint8_t* mask_host, int h, int w, uint8_t* out_host, size_t cap, size_t* out_len, int threads) {
  if (!mask_host || !out_host || h < 1 || w < 1) return 1;
  const size_t total = bd_png0_size(h, w);
  if (cap < total) return 1;
  crc_init();
  const uint64_t rowlen = static_cast<uint64_t>(w) + 1;
  const uint64_t n = static_cast<uint64_t>(h) * rowlen;
  const uint64_t nblk = std::max<uint64_t>(1, (n + BLOCK - 1) / BLOCK);
  uint8_t* o = out_host;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memcpy(o, sig, 8); o += 8;
  // IHDR
  put32be(o, 13); memcpy(o + 4, "IHDR", 4);
  put32be(o + 8, static_cast<uint32_t>(w)); put32be(o + 12, static_cast<uint32_t>(h));
  o[16] = 8; o[17] = 0; o[18] = 0; o[19] = 0; o[20] = 0;
  put32be(o + 21, crc32_update(0, o + 4, 17));
  o += 25;
  // IDAT: length, tag, zlib stream, crc
  const uint64_t zlen = 2 + n + 5 * nblk + 4;
  put32be(o, static_cast<uint32_t>(zlen)); memcpy(o + 4, "IDAT", 4);
  uint8_t* z = o + 8;
  z[0] = 0x78; z[1] = 0x01;
  uint8_t* body = z + 2;
  int T = threads > 0 ? threads : static_cast<int>(std::max(1u, std::min(32u, std::thread::hardware_concurrency())));
  T = static_cast<int>(std::min<uint64_t>(T, nblk));
  std::vector<uint32_t> crcs(T), adlers(T);
  std::vector<uint64_t> zbytes(T), rbytes(T);
  auto work = [&](int t) {
    const uint64_t k0 = nblk * t / T, k1 = nblk * (t + 1) / T;
    uint8_t* dst0 = body + k0 * (BLOCK + 5);
    uint8_t* dst = dst0;
    uint32_t ad = 1;
    uint64_t raw = 0;
    for (uint64_t k = k0; k < k1; ++k) {
      const uint64_t r0 = k * BLOCK, r1 = std::min(n, r0 + BLOCK), len = r1 - r0;
      dst[0] = (k + 1 == nblk) ? 1 : 0;
      dst[1] = len & 0xFF; dst[2] = len >> 8; dst[3] = ~len & 0xFF; dst[4] = (~len >> 8) & 0xFF;
      uint8_t* d = dst + 5;
      uint64_t r = r0;
      while (r < r1) {  // raw index -> (row, column); column 0 is the filter byte
        const uint64_t row = r / rowlen, col = r % rowlen;
        if (col == 0) { *d++ = 0; ++r; continue; }
        const uint64_t take = std::min(r1 - r, rowlen - col);
        memcpy(d, mask_host + row * static_cast<uint64_t>(w) + (col - 1), take);
        d += take; r += take;
      }
      ad = adler32_update(ad, dst + 5, len);
      raw += len;
      dst += 5 + len;
    }
    crcs[t] = crc32_update(0, dst0, static_cast<size_t>(dst - dst0));
    zbytes[t] = static_cast<uint64_t>(dst - dst0);
    adlers[t] = ad;
    rbytes[t] = raw;
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  uint32_t adler = 1;
  uint8_t* tail = body;
  for (int t = 0; t < T; ++t) { adler = t == 0 ? adlers[0] : adler32_combine(adler, adlers[t], rbytes[t]); tail += zbytes[t]; }
  put32be(tail, adler);
  // crc over "IDAT" + zlib stream: head (tag + 2 header bytes), the threads' parts, the Adler tail
  uint32_t crc = crc32_update(0, o + 4, 6);
  for (int t = 0; t < T; ++t) crc = crc32_combine(crc, crcs[t], zbytes[t]);
  crc = crc32_combine(crc, crc32_update(0, tail, 4), 4);
  put32be(tail + 4, crc);
  o = tail + 8;
  // IEND
  put32be(o, 0); memcpy(o + 4, "IEND", 4); put32be(o + 8, crc32_update(0, o + 4, 4));
  o += 12;
  if (out_len) *out_len = static_cast<size_t>(o - out_host);
  return (static_cast<size_t>(o - out_host) == total) ? 0 : 1;
}

}  // extern "C"
