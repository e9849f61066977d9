// Bitstream containers of the codec path, host-side C: the byte layouts a stream must have to be
// exchanged with the reference decoder.  No device work here; lives in the same C-ABI library.
//   * frame header       lossl_coord_int/model.py:447-452 (write), 466-473 (read)
//   * partition container lossl_coord_int/model.py:455-463 (write), 510-521 (read)
//   * BytesListUtils      lib/entropy_models/hyperprior/noisy_deep_factorized/utils.py:8-76
#include <string.h>

#include "common.cuh"

namespace {

inline int len_bytes_len(int64_t len) {  // math.ceil(len.bit_length() / 8) or 1
    int bits = 0;
    while (bits < 63 && (len >> bits) != 0) ++bits;
    const int b = (bits + 7) / 8;
    return b ? b : 1;
}
inline int bit_length(int v) {
    int b = 0;
    while (v >> b) ++b;
    return b;
}
// math.ceil(n / (8 // head_bits) + 0.25): head_bits 1 -> ceil((n + 2) / 8), head_bits 2 -> ceil((n + 1) / 4)
inline int64_t head_bytes_len(int n, int head_bits) {
    return head_bits == 1 ? ((int64_t)n + 2 + 7) / 8 : ((int64_t)n + 1 + 3) / 4;
}

}  // namespace

extern "C" int fpcc_frame_header_write(const int32_t coord_offset[3], int bottom_points, uint8_t out[8]) {
    FPCC_REQUIRE(coord_offset && out, "frame_header_write: NULL pointer");
    for (int a = 0; a < 3; ++a) {
        FPCC_REQUIRE(coord_offset[a] >= 0 && coord_offset[a] < 65536, "frame_header_write: coord_offset[%d]=%d does not fit 2 bytes", a, coord_offset[a]);
        out[2 * a] = (uint8_t)(coord_offset[a] & 0xff);
        out[2 * a + 1] = (uint8_t)(coord_offset[a] >> 8);
    }
    FPCC_REQUIRE(bottom_points >= 0 && bottom_points < 65536, "frame_header_write: bottom point count %d does not fit 2 bytes", bottom_points);
    out[6] = (uint8_t)(bottom_points & 0xff);
    out[7] = (uint8_t)(bottom_points >> 8);
    return FPCC_OK;
}

extern "C" int fpcc_frame_header_read(const uint8_t *data, int64_t len, int32_t coord_offset[3], int *bottom_points) {
    FPCC_REQUIRE(data && coord_offset && bottom_points, "frame_header_read: NULL pointer");
    FPCC_REQUIRE(len >= 8, "frame_header_read: stream of %lld bytes is shorter than its 8-byte header", (long long)len);
    for (int a = 0; a < 3; ++a) coord_offset[a] = (int32_t)data[2 * a] | ((int32_t)data[2 * a + 1] << 8);
    *bottom_points = (int)data[6] | ((int)data[7] << 8);
    return FPCC_OK;
}

// concat_bytes = b''.join(int_to_bytes(len(s), 3) + s)
extern "C" int64_t fpcc_partitions_pack(const uint8_t *const *items, const int64_t *lens, int n, uint8_t *out, int64_t out_cap) {
    if (!lens || n < 0 || (n > 0 && out && !items)) { fpcc::set_error("partitions_pack: bad arguments"); return -1; }
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (lens[i] < 0 || lens[i] >= ((int64_t)1 << 24)) { fpcc::set_error("partitions_pack: partition %d of %lld bytes does not fit a 3-byte length", i, (long long)lens[i]); return -1; }
        total += 3 + lens[i];
    }
    if (!out) return total;  // size query
    if (out_cap < total) { fpcc::set_error("partitions_pack: output buffer too small"); return -1; }
    uint8_t *p = out;
    for (int i = 0; i < n; ++i) {
        p[0] = (uint8_t)(lens[i] & 0xff); p[1] = (uint8_t)((lens[i] >> 8) & 0xff); p[2] = (uint8_t)((lens[i] >> 16) & 0xff);
        if (lens[i]) memcpy(p + 3, items[i], (size_t)lens[i]);
        p += 3 + lens[i];
    }
    return total;
}

// walks the container; offsets/lens may be NULL (count only).  Returns the partition count or -1 (truncated input).
extern "C" int fpcc_partitions_index(const uint8_t *data, int64_t len, int max_n, int64_t *offsets, int64_t *lens) {
    if (!data && len > 0) { fpcc::set_error("partitions_index: NULL pointer"); return -1; }
    int64_t pos = 0;
    int n = 0;
    while (pos != len) {
        if (len - pos < 3) { fpcc::set_error("partitions_index: truncated length prefix at byte %lld", (long long)pos); return -1; }
        const int64_t l = (int64_t)data[pos] | ((int64_t)data[pos + 1] << 8) | ((int64_t)data[pos + 2] << 16);
        pos += 3;
        if (len - pos < l) { fpcc::set_error("partitions_index: partition %d claims %lld bytes, %lld left", n, (long long)l, (long long)(len - pos)); return -1; }
        if (offsets && lens) {
            if (n >= max_n) { fpcc::set_error("partitions_index: more than %d partitions", max_n); return -1; }
            offsets[n] = pos; lens[n] = l;
        }
        pos += l;
        ++n;
    }
    return n;
}

// BytesListUtils.concat_bytes_list: head (marker bit, then len_bytes_len-1 of every item in head_bits bits each,
// right-aligned big endian; top bit of byte 0 set when head_bits == 2), the little-endian lengths, the payloads.
extern "C" int64_t fpcc_bytes_list_concat(const uint8_t *const *items, const int64_t *lens, int n, uint8_t *out, int64_t out_cap) {
    if (!lens || n < 2) { fpcc::set_error("bytes_list_concat: needs at least 2 items"); return -1; }
    int head_bits = 1;
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (lens[i] < 0) { fpcc::set_error("bytes_list_concat: negative length"); return -1; }
        const int lb = len_bytes_len(lens[i]);
        const int hb = bit_length(lb - 1);
        if (hb > head_bits) head_bits = hb;
        total += lb + lens[i];
    }
    if (head_bits > 2) { fpcc::set_error("bytes_list_concat: an item needs more than 4 length bytes"); return -1; }
    const int64_t hlen = head_bytes_len(n, head_bits);
    total += hlen;
    if (!out) return total;
    if (out_cap < total || !items) { fpcc::set_error("bytes_list_concat: output buffer too small"); return -1; }
    memset(out, 0, (size_t)hlen);
    const int64_t L = 1 + (int64_t)head_bits * n;  // bits used, right-aligned
    int64_t bit = hlen * 8 - L;                     // position from the MSB of byte 0
    auto put = [&](int v) { if (v) out[bit >> 3] |= (uint8_t)(0x80u >> (bit & 7)); ++bit; };
    put(1);
    for (int i = 0; i < n; ++i) {
        const int f = len_bytes_len(lens[i]) - 1;
        if (head_bits == 2) put((f >> 1) & 1);
        put(f & 1);
    }
    if (head_bits == 2) out[0] |= 0x80;
    uint8_t *p = out + hlen;
    for (int i = 0; i < n; ++i) {
        const int lb = len_bytes_len(lens[i]);
        for (int b = 0; b < lb; ++b) *p++ = (uint8_t)((lens[i] >> (8 * b)) & 0xff);
    }
    for (int i = 0; i < n; ++i) {
        if (lens[i]) memcpy(p, items[i], (size_t)lens[i]);
        p += lens[i];
    }
    return total;
}

// BytesListUtils.split_bytes_list: offsets/lens of the n items inside `data`; returns the bytes consumed or -1.
extern "C" int64_t fpcc_bytes_list_split(const uint8_t *data, int64_t len, int n, int64_t *offsets, int64_t *lens) {
    if (!data || !offsets || !lens || n < 1) { fpcc::set_error("bytes_list_split: bad arguments"); return -1; }
    if (len < 1) { fpcc::set_error("bytes_list_split: empty input"); return -1; }
    const int head_bits = (data[0] & 0x80) ? 2 : 1;
    const int64_t hlen = head_bytes_len(n, head_bits);
    if (len < hlen) { fpcc::set_error("bytes_list_split: truncated head"); return -1; }
    // the reference parses the head as an integer and drops everything up to and including the first set bit
    int64_t bit = 1;  // skip the flag bit of byte 0
    const int64_t nbits = hlen * 8;
    auto get = [&](int64_t b) { return (data[b >> 3] >> (7 - (b & 7))) & 1; };
    while (bit < nbits && !get(bit)) ++bit;
    if (bit >= nbits) { fpcc::set_error("bytes_list_split: head has no marker bit"); return -1; }
    ++bit;
    if (nbits - bit < (int64_t)head_bits * n) { fpcc::set_error("bytes_list_split: head too short for %d items", n); return -1; }
    int64_t pos = hlen;
    int64_t cursor_len = pos;
    for (int i = 0; i < n; ++i) {
        int f = 0;
        for (int b = 0; b < head_bits; ++b) f = (f << 1) | get(bit++);
        const int lb = f + 1;
        if (len - cursor_len < lb) { fpcc::set_error("bytes_list_split: truncated length table"); return -1; }
        int64_t l = 0;
        for (int b = 0; b < lb; ++b) l |= (int64_t)data[cursor_len + b] << (8 * b);
        lens[i] = l;
        cursor_len += lb;
    }
    pos = cursor_len;
    for (int i = 0; i < n; ++i) {
        if (len - pos < lens[i]) { fpcc::set_error("bytes_list_split: item %d claims %lld bytes, %lld left", i, (long long)lens[i], (long long)(len - pos)); return -1; }
        offsets[i] = pos;
        pos += lens[i];
    }
    return pos;
}
