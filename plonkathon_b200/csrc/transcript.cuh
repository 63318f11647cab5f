// Host-side Fiat-Shamir transcript: Merlin v1.0 over STROBE-128 / Keccak-f[1600], with the
// plonkathon message schedule on top.
//
// Replaces transcript.py:58-123 (`Transcript(MerlinTranscript)`) and the third-party `merlin` package
// the reference pins (poetry.lock:255-269; curdleproofs.pie @ 805d0678, absent from the reference tree):
// restated from the published Merlin / STROBE specifications.  Points and scalars are absorbed as 32-byte
// big-endian integers (transcript.py:62-67); a challenge is 255 squeezed bytes read as a big-endian integer
// mod r, redrawn while zero, and the raw bytes are then appended under the same label (transcript.py:69-75).
// This is host code (a few KiB of hashing per proof); it lives in the shared library so that a whole proof
// is one C-ABI call with no Python in the loop.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "field.cuh"

namespace pb200 {

inline uint64_t rol64(uint64_t x, int s) { return s ? (x << s) | (x >> (64 - s)) : x; }

inline void keccak_f1600(uint8_t* st) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
      0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
      0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61},
                                {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
  uint64_t A[5][5];
  for (int x = 0; x < 5; x++)
    for (int y = 0; y < 5; y++) {
      uint64_t v = 0;
      for (int k = 7; k >= 0; k--) v = (v << 8) | st[8 * (x + 5 * y) + k];
      A[x][y] = v;
    }
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t Cc[5], D[5], B[5][5];
    for (int x = 0; x < 5; x++) Cc[x] = A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4];
    for (int x = 0; x < 5; x++) D[x] = Cc[(x + 4) % 5] ^ rol64(Cc[(x + 1) % 5], 1);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) A[x][y] ^= D[x];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) B[y][(2 * x + 3 * y) % 5] = rol64(A[x][y], ROT[x][y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) A[x][y] = B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]);
    A[0][0] ^= RC[rnd];
  }
  for (int x = 0; x < 5; x++)
    for (int y = 0; y < 5; y++)
      for (int k = 0; k < 8; k++) st[8 * (x + 5 * y) + k] = (uint8_t)(A[x][y] >> (8 * k));
}

class Strobe128 {
 public:
  explicit Strobe128(const std::string& protocol_label) {
    memset(st_, 0, sizeof st_);
    const uint8_t init[6] = {1, kR + 2, 1, 0, 1, 96};
    memcpy(st_, init, 6);
    memcpy(st_ + 6, "STROBEv1.0.2", 12);
    keccak_f1600(st_);
    meta_ad((const uint8_t*)protocol_label.data(), protocol_label.size(), false);
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(kM | kA, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(kA, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n, bool more) { begin_op(kI | kA | kC, more); squeeze(out, n); }

 private:
  static const int kR = 166;
  static const uint8_t kI = 1, kA = 2, kC = 4, kT = 8, kM = 16, kK = 32;
  uint8_t st_[200];
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
  void run_f() {
    st_[pos_] ^= pos_begin_;
    st_[pos_ + 1] ^= 0x04;
    st_[kR + 1] ^= 0x80;
    keccak_f1600(st_);
    pos_ = 0;
    pos_begin_ = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      st_[pos_] ^= d[i];
      if (++pos_ == kR) run_f();
    }
  }
  void squeeze(uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      d[i] = st_[pos_];
      st_[pos_] = 0;
      if (++pos_ == kR) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;  // continuation of the current operation
    uint8_t old_begin = pos_begin_;
    pos_begin_ = pos_ + 1;
    cur_flags_ = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (kC | kK)) && pos_ != 0) run_f();
  }
};

class Transcript {
 public:
  explicit Transcript(const std::string& label) : strobe_("Merlin v1.0") {
    append_message("dom-sep", (const uint8_t*)label.data(), label.size());
  }
  void append_message(const std::string& label, const uint8_t* msg, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad((const uint8_t*)label.data(), label.size(), false);
    strobe_.meta_ad(len, 4, true);
    strobe_.ad(msg, n, false);
  }
  void challenge_bytes(const std::string& label, uint8_t* out, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad((const uint8_t*)label.data(), label.size(), false);
    strobe_.meta_ad(len, 4, true);
    strobe_.prf(out, n, false);
  }
  // transcript.py:62-63 -- canonical little-endian limbs in, 32-byte big-endian absorbed
  void append_scalar_le(const std::string& label, const uint8_t* le32) {
    uint8_t be[32];
    for (int i = 0; i < 32; i++) be[i] = le32[31 - i];
    append_message(label, be, 32);
  }
  // transcript.py:65-67 -- x then y under the same label
  void append_point_le(const std::string& label, const uint8_t* xy_le64) {
    append_scalar_le(label, xy_le64);
    append_scalar_le(label, xy_le64 + 32);
  }
  // transcript.py:69-75 -- returns the challenge as a canonical Fr (little-endian limbs)
  Fr get_and_append_challenge(const std::string& label) {
    for (;;) {
      uint8_t cb[255];
      challenge_bytes(label, cb, 255);
      Fr f = reduce_be(cb, 255);
      if (!f.is_zero()) {
        append_message(label, cb, 255);
        return f;
      }
    }
  }
  // big-endian byte string mod r, canonical form: Horner over 32-byte chunks, acc <- acc * 2^256 + chunk,
  // where multiplying by 2^256 is exactly a conversion to Montgomery form (R = 2^256)
  static Fr reduce_be(const uint8_t* b, size_t n) {
    Fr acc = Fr::zero();
    size_t pos = 0, first = n % 32 ? n % 32 : 32;
    while (pos < n) {
      size_t len = pos == 0 ? first : 32;
      Fr chunk = Fr::zero();  // up to 256 bits, may exceed r: fold the top bits first
      for (size_t k = 0; k < len; k++) {
        size_t bit = 8 * (len - 1 - k);
        chunk.v[bit >> 5] |= (uint32_t)b[pos + k] << (bit & 31);
      }
      // chunk < 2^256 < 6r: bring it below r with the Montgomery round trip (to_mont then from_mont reduce fully)
      chunk = fp_from_mont(fp_to_mont_any(chunk));
      acc = fp_add(fp_to_mont(acc), chunk);
      pos += len;
    }
    return acc;
  }

 private:
  Strobe128 strobe_;
};

}  // namespace pb200
