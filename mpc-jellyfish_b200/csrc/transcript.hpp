// transcript.hpp -- Fiat-Shamir transcripts of the PLONK prover, host side.
//
// Replaces plonk/src/transcript/solidity.rs:31-78 (`SolidityTranscript`: Keccak-256, what the
// collaborative prover uses through mpc_transcript.rs:30-52) and plonk/src/transcript/standard.rs:18-46
// (`StandardTranscript`: merlin 3 / STROBE-128 over Keccak-f[1600]).  The sha3 and merlin crates
// are un-vendored dependencies of the reference; their published algorithms are restated here
// and pinned by the reference's own Keccak known-answer test (solidity.rs:80-96) and merlin's
// published test vector (tests/test_host_transcript.py).
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

namespace jf {

inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

// Keccak-f[1600] on 25 little-endian lanes, lane (x, y) at index x + 5 y.
inline void keccak_f1600(uint64_t s[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
        0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
        0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
        0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
        0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    // rho offsets walked along the pi permutation cycle starting from lane 1
    static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t c[5];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) s[y + x] ^= d;
        }
        uint64_t cur = s[1];
        for (int i = 0; i < 24; i++) {
            uint64_t nxt = s[PI[i]];
            s[PI[i]] = rotl64(cur, RHO[i]);
            cur = nxt;
        }
        for (int y = 0; y < 25; y += 5) {
            uint64_t r[5];
            for (int x = 0; x < 5; x++) r[x] = s[y + x];
            for (int x = 0; x < 5; x++) s[y + x] = r[x] ^ (~r[(x + 1) % 5] & r[(x + 2) % 5]);
        }
        s[0] ^= RC[round];
    }
}

// byte view helpers (host is little-endian: x86-64 / aarch64)
inline uint8_t *lane_bytes(uint64_t *s) { return reinterpret_cast<uint8_t *>(s); }

// Keccak-256 as in the sha3 crate's `Keccak256` (pre-NIST padding 0x01 ... 0x80, rate 136).
inline void keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t s[25];
    memset(s, 0, sizeof s);
    uint8_t *b = lane_bytes(s);
    while (len >= rate) {
        for (size_t i = 0; i < rate; i++) b[i] ^= data[i];
        keccak_f1600(s);
        data += rate;
        len -= rate;
    }
    for (size_t i = 0; i < len; i++) b[i] ^= data[i];
    b[len] ^= 0x01;
    b[rate - 1] ^= 0x80;
    keccak_f1600(s);
    memcpy(out, b, 32);
}

// STROBE-128 as instantiated by merlin (strobe.rs of the merlin crate)
class Strobe128 {
  public:
    explicit Strobe128(const std::string &protocol_label) {
        memset(st_, 0, sizeof st_);
        uint8_t *b = lane_bytes(st_);
        const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(b, init, 6);
        memcpy(b + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st_);
        meta_ad(reinterpret_cast<const uint8_t *>(protocol_label.data()), protocol_label.size(), false);
    }
    void meta_ad(const uint8_t *d, size_t n, bool more) {
        begin_op(FLAG_M | FLAG_A, more);
        absorb(d, n);
    }
    void ad(const uint8_t *d, size_t n, bool more) {
        begin_op(FLAG_A, more);
        absorb(d, n);
    }
    void prf(uint8_t *out, size_t n, bool more) {
        begin_op(FLAG_I | FLAG_A | FLAG_C, more);
        squeeze(out, n);
    }

  private:
    static constexpr uint8_t R = 166;
    static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
    uint64_t st_[25];
    uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;

    void run_f() {
        uint8_t *b = lane_bytes(st_);
        b[pos_] ^= pos_begin_;
        b[pos_ + 1] ^= 0x04;
        b[R + 1] ^= 0x80;
        keccak_f1600(st_);
        pos_ = 0;
        pos_begin_ = 0;
    }
    void absorb(const uint8_t *d, size_t n) {
        uint8_t *b = lane_bytes(st_);
        for (size_t i = 0; i < n; i++) {
            b[pos_] ^= d[i];
            if (++pos_ == R) run_f();
        }
    }
    void squeeze(uint8_t *out, size_t n) {
        uint8_t *b = lane_bytes(st_);
        for (size_t i = 0; i < n; i++) {
            out[i] = b[pos_];
            b[pos_] = 0;
            if (++pos_ == R) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;  // continuation of the same operation
        uint8_t old_begin = pos_begin_;
        pos_begin_ = pos_ + 1;
        cur_flags_ = flags;
        const uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
    }
};

// `PlonkTranscript` (plonk/src/transcript/mod.rs:40-215): append_message + 64-byte challenge source.
class Transcript {
  public:
    enum Kind { SOLIDITY = 0, STANDARD = 1 };
    Transcript(int kind, const std::string &label) : kind_(kind), strobe_("Merlin v1.0") {
        memset(state_, 0, sizeof state_);
        if (kind_ == STANDARD) merlin_append("dom-sep", reinterpret_cast<const uint8_t *>(label.data()), label.size());
    }
    void append_message(const char *label, const uint8_t *msg, size_t n) {
        if (kind_ == SOLIDITY) buf_.insert(buf_.end(), msg, msg + n);  // labels ignored (solidity.rs:43-47)
        else merlin_append(label, msg, n);
    }
    // Bytes the challenge is reduced from: state[..48] (solidity.rs:53-77) or 64 PRF bytes
    // (standard.rs:33-45).  Returns the count; the caller reduces mod r and, for STANDARD,
    // re-appends the canonical challenge through append_message (done in challenge_done).
    size_t challenge_bytes(const char *label, uint8_t out[64]) {
        if (kind_ == SOLIDITY) {
            std::vector<uint8_t> in(64 + buf_.size() + 1);
            memcpy(in.data(), state_, 64);
            if (!buf_.empty()) memcpy(in.data() + 64, buf_.data(), buf_.size());
            in.back() = 0;
            keccak256(in.data(), in.size(), state_);
            uint8_t second[32];
            in.back() = 1;
            keccak256(in.data(), in.size(), second);
            memcpy(state_ + 32, second, 32);
            memcpy(out, state_, 48);
            return 48;
        }
        const uint32_t n = 64;
        const size_t ll = strlen(label);
        strobe_.meta_ad(reinterpret_cast<const uint8_t *>(label), ll, false);
        strobe_.meta_ad(reinterpret_cast<const uint8_t *>(&n), 4, true);
        strobe_.prf(out, 64, false);
        return 64;
    }
    // StandardTranscript appends the reduced challenge (32 canonical LE bytes) under the same label.
    void challenge_done(const char *label, const uint8_t *canonical, size_t n) {
        if (kind_ == STANDARD) merlin_append(label, canonical, n);
    }

  private:
    int kind_;
    std::vector<uint8_t> buf_;
    uint8_t state_[64];
    Strobe128 strobe_;
    void merlin_append(const char *label, const uint8_t *msg, size_t n) {
        const uint32_t len = (uint32_t)n;
        strobe_.meta_ad(reinterpret_cast<const uint8_t *>(label), strlen(label), false);
        strobe_.meta_ad(reinterpret_cast<const uint8_t *>(&len), 4, true);
        strobe_.ad(msg, n, false);
    }
};

}  // namespace jf
