/* crc_emul.cpp -- CPU replay of k_music_crc's decomposition (lamejs_b200/csrc/k_tag.cuh compiled by g++: the header's
 * __host__ __device__ functions are everything that decides a bit; the kernel adds the warp xor and the atomic xor, which
 * are replayed here as plain xors in piece / lane order and in reverse order).  TEST INFRASTRUCTURE (tests/test_tag_cpu.py). */
#include "../lamejs_b200/csrc/k_tag.cuh"

extern "C" {

void emul_tables(CrcTables* t) { crc_host_tables(t); }

/* the kernel's result for one byte range; order != 0 walks pieces and lanes backwards (xor must not care) */
unsigned emul_range_crc(const uint8_t* buf, long long len, int order) {
  static CrcTables t;
  static bool ready = false;
  if (!ready) { crc_host_tables(&t); ready = true; }
  unsigned acc = 0;
  const long long pieces = (len + CRC_PIECE_BYTES - 1) / CRC_PIECE_BYTES;
  for (long long q = 0; q < pieces; q++) {
    const long long piece = order ? pieces - 1 - q : q;
    unsigned c = 0;
    long long after_piece = 0;
    for (int l = 0; l < 32; l++) {
      const int lane = order ? 31 - l : l;
      const CrcLanePlan p = crc_plan(len, piece, lane);
      c ^= crc_lane(buf, p, t.byte_table, t.pow);
      after_piece = p.after_piece;
    }
    acc ^= crc_shift(c, (unsigned long long)after_piece, t.pow);
  }
  return acc;
}

unsigned emul_append(unsigned crc_old, unsigned crc_new, unsigned long long nbytes) {
  static CrcTables t;
  static bool ready = false;
  if (!ready) { crc_host_tables(&t); ready = true; }
  return crc_append(crc_old, crc_new, nbytes, t.pow);
}

unsigned emul_shift(unsigned crc, unsigned long long nbytes) {
  static CrcTables t;
  static bool ready = false;
  if (!ready) { crc_host_tables(&t); ready = true; }
  return crc_shift(crc, nbytes, t.pow);
}

}
