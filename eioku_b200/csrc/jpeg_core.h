// jpeg_core.h -- baseline JPEG (ISO/IEC 10918-1, 8-bit, Huffman, 4:2:0) decoding arithmetic shared by the CUDA kernels of
// libesd_decode.so (csrc/esd_decode.cu) and by a CPU test shim (tests/jpeg_shim.cpp), so that the exact same code is checked
// on the CPU against cv2.imdecode before it runs on a GPU.
//
// The arithmetic restates libjpeg's defaults, which is what OpenCV's imdecode runs (libjpeg-turbo: dct_method JDCT_ISLOW,
// do_fancy_upsampling TRUE; its SIMD paths are bit-exact with the C code):
//   * jidctint.c  jpeg_idct_islow      -- Loeffler-Ligtenberg-Moschytz, CONST_BITS 13, PASS1_BITS 2
//   * jdsample.c  h2v2_fancy_upsample  -- triangle filter, 3/4 - 1/4 vertically then horizontally, bias 8 / 7
//   * jdcolor.c   ycc_rgb_convert      -- 16-bit fixed-point JFIF YCbCr -> RGB
// so a decoded picture is bit-identical to cv2.imdecode's (tests/test_host_jpeg.py on the CPU, tests/test_gpu_decode.py on
// the device).  Not a general JPEG decoder: progressive / arithmetic / 12-bit / other subsamplings are rejected by the host
// parser and those files go to nvJPEG or to the host decoder.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define JPG_HD __host__ __device__ __forceinline__
#define JPG_M __host__ __device__ __forceinline__   // member functions
#else
#define JPG_HD static inline
#define JPG_M inline
#endif

namespace esdjpeg {

constexpr int kLookBits = 9;  // Huffman look-ahead table width

// zig-zag position -> natural (row-major) position inside an 8x8 block (device code keeps its own __constant__ copy and
// passes it to decode_block)
#define ESD_JPEG_NATURAL_ORDER                                                                                                   \
    {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, \
     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}
static const uint8_t kNaturalOrderHost[64] = ESD_JPEG_NATURAL_ORDER;

// One Huffman table in decoding form (libjpeg's d_derived_tbl, reduced).
struct HuffTable {
    uint16_t look[1 << kLookBits];  // (code length << 8) | symbol for codes of <= kLookBits bits, 0 = longer code
    int32_t maxcode[18];            // largest code of each length (-1 = none); [17] = sentinel
    int32_t valoffset[17];          // huffval index of the first code of each length, minus that code
    uint32_t limit[17];             // left-aligned 16-bit windows below limit[L] hold a code of at most L bits (canonical codes)
    uint8_t huffval[256];
};

// Everything one scan needs (per file; per-frame quantisation tables are allowed to differ and travel separately).
struct ScanTables {
    HuffTable dc[2], ac[2];
};

struct FrameGeometry {
    int width, height;          // picture
    int mcus_x, mcus_y;         // 16x16 MCUs
    int yblocks_x, cblocks_x;   // blocks per row: luma 2 * mcus_x, chroma mcus_x
    int restart_interval;       // MCUs between RSTn markers, 0 = none
};

// Builds `t` from the DHT payload (16 counts + values).  Returns false for an invalid table.
JPG_HD bool build_huff_table(const uint8_t* counts, const uint8_t* vals, int nvals, HuffTable* t) {
    int total = 0;
    for (int i = 0; i < 16; ++i) total += counts[i];
    if (total > 256 || total != nvals) return false;
    for (int i = 0; i < 256; ++i) t->huffval[i] = i < nvals ? vals[i] : 0;
    for (int i = 0; i < (1 << kLookBits); ++i) t->look[i] = 0;
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        const int n = counts[len - 1];
        t->valoffset[len] = k - code;
        if (n) {
            for (int j = 0; j < n; ++j) {
                if (len <= kLookBits) {
                    const int first = (code + j) << (kLookBits - len);
                    for (int f = 0; f < (1 << (kLookBits - len)); ++f) t->look[first + f] = (uint16_t)((len << 8) | t->huffval[k + j]);
                }
            }
            code += n;
            k += n;
            t->maxcode[len] = code - 1;
            if (code > (1 << len)) return false;
        } else {
            t->maxcode[len] = -1;
        }
        t->limit[len] = (uint32_t)code << (16 - len);
        code <<= 1;
    }
    t->maxcode[0] = -1;
    t->valoffset[0] = 0;
    t->limit[0] = 0;
    t->maxcode[17] = 0x7fffffff;
    return true;
}

// MSB-first bit reader over an entropy-coded segment with byte stuffing (FF00 -> FF).  A marker ends the data: zeros follow.
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;  // valid bits are the top `bits`
    int bits;
    int marker;    // marker code seen (0 = none)

    JPG_M void init(const uint8_t* data, int len) { p = data; end = data + len; buf = 0; bits = 0; marker = 0; }

    JPG_M void fill() {
        // fast path: four bytes without an FF
        if (bits <= 32 && !marker && p + 4 <= end) {
            const uint32_t w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
            const uint32_t v = ~w;
            if (!((v - 0x01010101u) & ~v & 0x80808080u)) {  // no byte of w is 0xFF
                buf |= (uint64_t)w << (32 - bits);
                bits += 32;
                p += 4;
                return;
            }
        }
        while (bits <= 56) {
            uint32_t b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    while (p < end && *p == 0xFF) ++p;        // fill bytes before a marker
                    const uint32_t b2 = p < end ? *p : 0xD9u;
                    if (b2 == 0) ++p;                          // stuffed zero: the data byte 0xFF
                    else { marker = (int)b2; --p; b = 0; }     // a marker ends the data (p stays on its 0xFF); zeros follow
                }
            }
            buf |= (uint64_t)b << (56 - bits);
            bits += 8;
        }
    }
    JPG_M uint32_t peek16() {
        if (bits < 16) fill();
        return (uint32_t)(buf >> 48);
    }
    JPG_M void skip(int n) { buf <<= n; bits -= n; }
    JPG_M int32_t receive_extend(int s) {  // s in 1..15 (16 for the degenerate DC case)
        if (bits < s) fill();
        const uint32_t v = (uint32_t)(buf >> (64 - s));
        skip(s);
        // HUFF_EXTEND: values below 2^(s-1) are negative
        return (int32_t)v < (1 << (s - 1)) ? (int32_t)v - (1 << s) + 1 : (int32_t)v;
    }
    // after a restart interval: drop the partial byte, step over the RSTn marker
    JPG_M void restart() {
        buf = 0; bits = 0;
        if (marker >= 0xD0 && marker <= 0xD7) { p += 2; marker = 0; }
        else if (!marker) {  // marker not reached through the bit buffer yet: find it
            while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
            if (p + 1 < end) p += 2;
        }
    }
};

JPG_HD int decode_symbol(BitReader& br, const HuffTable& t) {
    const uint32_t pk = br.peek16();
    const uint32_t e = t.look[pk >> (16 - kLookBits)];
    if (e) {
        br.skip((int)(e >> 8));
        return (int)(e & 255u);
    }
    int len = kLookBits + 1;
    int32_t code = (int32_t)(pk >> (16 - len));
    while (len <= 16 && code > t.maxcode[len]) {
        ++len;
        code = (int32_t)(pk >> (16 - len));
    }
    if (len > 16) { br.skip(16); return 0; }  // corrupt data: resynchronise on garbage rather than loop
    br.skip(len);
    return t.huffval[(code + t.valoffset[len]) & 255];
}

// One 8x8 block: DC difference + AC run/size pairs -> coefficients in NATURAL order (the caller has zeroed `coef`).
// Returns the zig-zag index of the last non-zero coefficient.
JPG_HD int decode_block(BitReader& br, const HuffTable& dc, const HuffTable& ac, const uint8_t* natural, int& pred, int16_t* coef) {
    int s = decode_symbol(br, dc);
    if (s) pred += br.receive_extend(s > 16 ? 16 : s);
    coef[0] = (int16_t)pred;
    int last = 0;
    for (int k = 1; k < 64;) {
        const int rs = decode_symbol(br, ac);
        const int r = rs >> 4;
        s = rs & 15;
        if (s == 0) {
            if (r != 15) break;  // EOB
            k += 16;
            continue;
        }
        k += r;
        if (k > 63) break;  // corrupt
        coef[natural[k]] = (int16_t)br.receive_extend(s);
        last = k;
        ++k;
    }
    return last;
}


// ---- scan decoding as ONE flat loop ---------------------------------------------------------------------------------------
// The same decoding as decode_block over a whole interleaved 4:2:0 scan, written so that every symbol takes the same
// instruction path: one iteration = one Huffman symbol (a DC size or an AC run/size) plus its extra bits, the position inside
// the block / MCU / picture is data, not control flow.  On the GPU a warp decodes 32 PICTURES in lock-step this way (the
// nested-loop form diverges at every branch and ran ~1 500 cycles per symbol; profiles/r02_decode_probe_native_v1.log).
// Input: the scan with the byte stuffing already removed (FF 00 -> FF), starting on a 4-byte boundary, `nwords` 32-bit words
// long including at least two words of zero padding, no restart markers.  `cf` receives the non-zero coefficients of the
// picture (natural order inside a block; blocks in decoding order) and has room for ONE SPARE BLOCK behind the last (scratch);
// the caller has zeroed it.
JPG_HD void decode_scan_flat(const uint32_t* __restrict__ words, int nwords, const ScanTables& T, const int* td, const int* ta,
                             const uint8_t* __restrict__ natural, int n_mcus, int16_t* __restrict__ cf) {
    // A single warp per SM runs this (the per-picture Huffman tables fill the shared memory), so the loop is bound by the
    // LATENCY of its instruction stream (~4 cycles per instruction for a lone warp), not by issue slots.  Hence: (1) selects
    // and arithmetic instead of branches (first version: 192 warp instructions per symbol with 12.7 of 32 lanes active);
    // (2) coefficients are stored in DECODING order (block s of the scan at cf + 64 s: Y00 Y01 Y10 Y11 Cb Cr per MCU), so the
    // loop knows nothing about the picture's geometry -- the IDCT kernel, which is parallel, maps blocks to positions;
    // (3) the table selectors of the next block are computed in the shadow of the look-up and merely selected when a block
    // ends; (4) code and extra bits are taken from one 32-bit window, one 64-bit shift per symbol; (5) the next input word is
    // loaded one refill ahead.
    const HuffTable* const tab0 = &T.dc[0];              // dc[0], dc[1], ac[0], ac[1] are contiguous
    const uint32_t dpack = (uint32_t)td[0] | ((uint32_t)td[1] << 8) | ((uint32_t)td[2] << 16);                    // table selector of
    const uint32_t apack = (uint32_t)(2 + ta[0]) | ((uint32_t)(2 + ta[1]) << 8) | ((uint32_t)(2 + ta[2]) << 16);  // component c in byte c
    const int dsel0 = td[0], asel0 = 2 + ta[0];
    uint64_t buf = 0;
    int bits = 0, wi = 1;
    uint32_t nextw = nwords > 0 ? words[0] : 0u;         // loaded one refill ahead
    int pred0 = 0, pred1 = 0, pred2 = 0;
    int b = 0, comp = 0, kpos = 0;
    int blk = 0;                                         // coefficient offset of the current block
    int dsel = dsel0, asel = asel0;
    int remaining = 6 * n_mcus;
    const int spare = 6 * n_mcus * 64;                   // one spare block behind the last: the sink of coefficient-less symbols
    while (remaining > 0) {
        // ---- selectors of the next block (independent of the symbols decoded below)
        const int nb = b == 5 ? 0 : b + 1;
        const int ncomp = nb < 4 ? 0 : nb - 3;
        const int ndsel = (int)((dpack >> (8 * ncomp)) & 255u);
        const int nasel = (int)((apack >> (8 * ncomp)) & 255u);
        // ---- refill: at most 16 + 15 bits are consumed per symbol
        const bool need = bits <= 32;
        const uint32_t nm = need ? 0xffffffffu : 0u;     // masks instead of ?: below -- the compiler turned the selects into divergent branches
        uint32_t w = nextw;
        w = (w >> 24) | ((w >> 8) & 0xff00u) | ((w << 8) & 0xff0000u) | (w << 24);  // big-endian bit order
        buf |= ((uint64_t)(w & nm) << ((32 - bits) & 63));
        bits += (int)(32u & nm);
        const uint32_t fetched = words[wi < nwords ? wi : nwords - 1];  // unconditional (same word until consumed: it stays in L1)
        nextw = (fetched & nm) | (nextw & ~nm);
        wi += (int)(1u & nm);
        // ---- one symbol
        const uint32_t win = (uint32_t)(buf >> 32);
        const uint32_t pk = win >> 16;
        const bool is_dc = kpos == 0;
        const HuffTable* tab = tab0 + (is_dc ? dsel : asel);
        const uint32_t e = tab->look[pk >> (16 - kLookBits)];
        int len = (int)(e >> 8), sym = (int)(e & 255u);
        if (e == 0) {  // a code longer than the look-ahead: its length from the canonical limits, no loop
            len = kLookBits + 1;
#pragma unroll
            for (int L = kLookBits + 1; L < 16; ++L) len += pk >= tab->limit[L] ? 1 : 0;
            sym = tab->huffval[((int)(pk >> (16 - len)) + tab->valoffset[len]) & 255];
        }
        const int dm = is_dc ? -1 : 0;                                     // all ones for the DC symbol of a block
        const int size = ((sym > 15 ? 15 : sym) & dm) | ((sym & 15) & ~dm);
        const int run = (sym >> 4) & ~dm;
        const int v = (int)(((win << len) >> 1) >> (31 - size));          // the `size` bits behind the code (0 for size == 0)
        buf <<= len + size;
        bits -= len + size;
        const int neg = v < ((1 << size) >> 1) ? -1 : 0;                   // HUFF_EXTEND; size == 0 gives 0
        const int val = v + ((1 - (1 << size)) & neg);
        // DC: accumulate the prediction of this component
        pred0 += val & dm & (comp == 0 ? -1 : 0);
        pred1 += val & dm & (comp == 1 ? -1 : 0);
        pred2 += val & dm & (comp == 2 ? -1 : 0);
        const int pred = (pred0 & (comp == 0 ? -1 : 0)) | (pred1 & (comp == 1 ? -1 : 0)) | (pred2 & (comp == 2 ? -1 : 0));
        const int kk = kpos + run;                                       // zig-zag position of an AC coefficient
        const int cm = (is_dc || (size != 0 && kk < 64)) ? -1 : 0;       // a coefficient is stored
        const int where = (int)natural[kk & 63] & ~dm;
        // unconditional store: symbols that carry no coefficient (EOB, ZRL) write to the spare block behind the picture's last
        cf[((blk + where) & cm) | (spare & ~cm)] = (int16_t)((pred & dm) | (val & ~dm));
        const int zm = size != 0 ? -1 : 0;
        const int k_ac = ((kk + 1) & zm) | ((run == 15 ? kpos + 16 : 64) & ~zm);
        kpos = (1 & dm) | (k_ac & ~dm);
        // ---- end of block
        const int em = kpos >= 64 ? -1 : 0;
        remaining += em;
        kpos &= ~em;
        blk += 64 & em;
        b = (nb & em) | (b & ~em);
        comp = (ncomp & em) | (comp & ~em);
        dsel = (ndsel & em) | (dsel & ~em);
        asel = (nasel & em) | (asel & ~em);
    }
}

// ---- one picture decoded by MANY threads: speculative sub-sequences that self-synchronise -----------------------------------
// A scan without restart markers is one dependent chain -- yet Huffman streams resynchronise: a decoder started at an arbitrary bit
// with a guessed state (block-in-MCU b, zig-zag position k) produces garbage for a while and then, with overwhelming probability,
// falls into step with the true decoding and stays there (Klein & Wiseman; used for JPEG by Weissenberger & Schmidt, "Accelerating
// JPEG decompression on GPUs", 2021).  The scan is cut into sub-sequences of S bits, thread t owns the symbols that START in
// [t S, (t+1) S).  Round 0: every thread decodes its sub-sequence from the guess (p = t S, b = 0, k = 0) -- thread 0's start is the
// true one.  Round r: thread t restarts from the END state thread t-1 reached in round r-1, if that differs from what it started
// from before.  A fixed point is the sequential decoding (correct states spread at least one sub-sequence per round, in practice
// they are there after 2-4 rounds).  Then block counts and DC sums are prefix-summed over the threads and a last pass writes the
// coefficients -- DC absolute -- exactly where decode_scan_flat puts them.
struct SpanState {
    uint32_t p;   // bit position (in the unstuffed scan) where the next symbol starts
    uint16_t b;   // block inside the MCU, 0..5 (0-3 luma, 4 Cb, 5 Cr)
    uint16_t k;   // zig-zag position of the next coefficient; 0 = the next symbol is the DC size of a block
};
struct SpanResult {
    SpanState end;     // state after the last symbol that starts before the limit
    int n_blocks;      // blocks COMPLETED by those symbols
    int dc[3];         // sum of the DC differences decoded, per component
    int n_coefs;       // coefficients those symbols carry (DC of every block started + non-zero AC): entries of the sparse hand-off
};
// decode_span modes
enum { SPAN_COUNT = 0, SPAN_DENSE = 1, SPAN_SPARSE = 2 };
// Sparse hand-off to the IDCT (SPAN_SPARSE): instead of scattering 16-bit coefficients into a cleared [blocks][64] array, the
// symbols are appended to a list -- entry = value (low 16 bits) | natural position << 16, a block's entries contiguous, its DC
// first -- and `bstart[blk]` records where block blk's entries begin (bstart[total_blocks] closes the last block).  A 1080p
// picture moves ~1.3 MB this way instead of 6.3 MB cleared + 6.3 MB read back.  List capacity `cap` entries (+ 1 sink entry at
// [cap]); bstart has total_blocks + 2 entries (the last is a sink).
JPG_HD bool same_state(const SpanState& a, const SpanState& b) { return a.p == b.p && a.b == b.b && a.k == b.k; }

// Decodes the symbols that start in [start.p, limit_bit).  WRITE: coefficients go to cf (decoding order, natural order inside a
// block) with `blk0` the index of the block in progress at the start, `pred` the DC predictors there, blocks >= total_blocks (the
// zero padding behind the last real block decodes as garbage) sent to the spare block at cf + 64 * total_blocks.
template <int MODE>
JPG_HD SpanResult decode_span(const uint32_t* __restrict__ words, int nwords, const ScanTables& T, const int* td, const int* ta,
                             const uint8_t* __restrict__ natural, SpanState start, uint32_t limit_bit, int16_t* __restrict__ cf,
                             int blk0, const int* pred_in, int total_blocks, uint32_t* __restrict__ bstart = nullptr, int o0 = 0,
                             int cap = 0) {
    constexpr bool WRITE = MODE != SPAN_COUNT;
    uint32_t* const list = reinterpret_cast<uint32_t*>(cf);   // SPAN_SPARSE: the same buffer holds the entry list
    int o = o0;
    const HuffTable* const tab0 = &T.dc[0];
    const uint32_t dpack = (uint32_t)td[0] | ((uint32_t)td[1] << 8) | ((uint32_t)td[2] << 16);
    const uint32_t apack = (uint32_t)(2 + ta[0]) | ((uint32_t)(2 + ta[1]) << 8) | ((uint32_t)(2 + ta[2]) << 16);
    uint32_t p = start.p;
    int b = (int)start.b, kpos = (int)start.k;
    int comp = b < 4 ? 0 : b - 3;
    int dsel = (int)((dpack >> (8 * comp)) & 255u), asel = (int)((apack >> (8 * comp)) & 255u);
    SpanResult R;
    R.n_blocks = 0;
    R.n_coefs = 0;
    R.dc[0] = R.dc[1] = R.dc[2] = 0;
    int pred0 = WRITE ? pred_in[0] : 0, pred1 = WRITE ? pred_in[1] : 0, pred2 = WRITE ? pred_in[2] : 0;
    int blk = blk0;
    const int last = nwords - 1;
    auto be = [](uint32_t w) { return (w >> 24) | ((w >> 8) & 0xff00u) | ((w << 8) & 0xff0000u) | (w << 24); };
    // bit buffer positioned at p: the top `bits` bits of buf are the stream from p on
    int wi = (int)(p >> 5);
    const int r0 = (int)(p & 31u);
    uint64_t buf = (((uint64_t)be(words[wi < last ? wi : last]) << 32) | (uint64_t)be(words[wi + 1 < last ? wi + 1 : last])) << r0;
    int bits = 64 - r0;
    wi += 2;
    uint32_t nextw = words[wi < last ? wi : last];
    while (p < limit_bit) {
        const int nb = b == 5 ? 0 : b + 1;
        const int ncomp = nb < 4 ? 0 : nb - 3;
        const int ndsel = (int)((dpack >> (8 * ncomp)) & 255u);
        const int nasel = (int)((apack >> (8 * ncomp)) & 255u);
        const uint32_t nm = bits <= 32 ? 0xffffffffu : 0u;
        buf |= ((uint64_t)(be(nextw) & nm) << ((32 - bits) & 63));
        bits += (int)(32u & nm);
        wi += (int)(1u & nm);
        const uint32_t fetched = words[wi < last ? wi : last];
        nextw = (fetched & nm) | (nextw & ~nm);
        const uint32_t win = (uint32_t)(buf >> 32);
        const uint32_t pk = win >> 16;
        const bool is_dc = kpos == 0;
        const HuffTable* tab = tab0 + (is_dc ? dsel : asel);
        const uint32_t e = tab->look[pk >> (16 - kLookBits)];
        int len = (int)(e >> 8), sym = (int)(e & 255u);
        if (e == 0) {
            len = kLookBits + 1;
#pragma unroll
            for (int L = kLookBits + 1; L < 16; ++L) len += pk >= tab->limit[L] ? 1 : 0;
            sym = tab->huffval[((int)(pk >> (16 - len)) + tab->valoffset[len]) & 255];
        }
        const int dm = is_dc ? -1 : 0;
        const int size = ((sym > 15 ? 15 : sym) & dm) | ((sym & 15) & ~dm);
        const int run = (sym >> 4) & ~dm;
        const int v = (int)(((win << len) >> 1) >> (31 - size));
        buf <<= len + size;
        bits -= len + size;
        p += (uint32_t)(len + size);
        const int neg = v < ((1 << size) >> 1) ? -1 : 0;
        const int val = v + ((1 - (1 << size)) & neg);
        const int c0 = comp == 0 ? -1 : 0, c1 = comp == 1 ? -1 : 0, c2 = comp == 2 ? -1 : 0;
        R.dc[0] += val & dm & c0;
        R.dc[1] += val & dm & c1;
        R.dc[2] += val & dm & c2;
        const int kk = kpos + run;
        const int carries = (is_dc || (size != 0 && kk < 64)) ? -1 : 0;   // this symbol carries a coefficient
        R.n_coefs -= carries;
        if (WRITE) {
            pred0 += val & dm & c0;
            pred1 += val & dm & c1;
            pred2 += val & dm & c2;
            const int pred = (pred0 & c0) | (pred1 & c1) | (pred2 & c2);
            const int where = (int)natural[kk & 63] & ~dm;
            const int value = (pred & dm) | (val & ~dm);
            if (MODE == SPAN_DENSE) {
                const int cm = carries & (blk < total_blocks ? -1 : 0);
                cf[((blk * 64 + where) & cm) | ((total_blocks * 64) & ~cm)] = (int16_t)value;
            } else {
                // where block blk's entries begin (its DC symbol comes first); blocks past the picture's end go to the sink slot
                const int bi = (is_dc && blk <= total_blocks) ? blk : total_blocks + 1;
                bstart[bi] = (uint32_t)o;
                const int ok = carries & ((blk < total_blocks && o < cap) ? -1 : 0);
                list[(o & ok) | (cap & ~ok)] = ((uint32_t)value & 0xffffu) | ((uint32_t)where << 16);
                o -= carries;
            }
        }
        const int zm = size != 0 ? -1 : 0;
        const int k_ac = ((kk + 1) & zm) | ((run == 15 ? kpos + 16 : 64) & ~zm);
        kpos = (1 & dm) | (k_ac & ~dm);
        const int em = kpos >= 64 ? -1 : 0;
        R.n_blocks -= em;
        kpos &= ~em;
        blk -= em;
        b = (nb & em) | (b & ~em);
        comp = (ncomp & em) | (comp & ~em);
        dsel = (ndsel & em) | (dsel & ~em);
        asel = (nasel & em) | (asel & ~em);
    }
    R.end.p = p;
    R.end.b = (uint16_t)b;
    R.end.k = (uint16_t)kpos;
    // the picture's last block ended exactly at this span's end and nothing started behind it: close it here
    if (MODE == SPAN_SPARSE && blk == total_blocks && kpos == 0) bstart[total_blocks] = (uint32_t)o;
    return R;
}

// Sparse hand-off, consumer side: block s of a picture as 64 coefficients (natural order).  Entries are clamped to the list and to
// 64 per block, so stale or corrupt offsets can produce a wrong block but no out-of-bounds access.
JPG_HD void expand_block(const uint32_t* __restrict__ list, const uint32_t* __restrict__ bstart, int s, int cap, int16_t* coef64) {
    for (int i = 0; i < 64; ++i) coef64[i] = 0;
    uint32_t en = bstart[s + 1], st = bstart[s];
    if (en > (uint32_t)cap) en = (uint32_t)cap;
    if (st > en) st = en;
    if (en - st > 64u) en = st + 64u;
    for (uint32_t e = st; e < en; ++e) coef64[(list[e] >> 16) & 63u] = (int16_t)(list[e] & 0xffffu);
}

// Sub-sequence length for a scan of `nwords` words and `total_blocks` blocks decoded by at most `max_threads` threads: a multiple
// of 32 bits, >= 512, and long enough to hold ~32 blocks -- states fall into step at block boundaries, and with only a few
// boundaries per sub-sequence (large blocks: noisy content at high quality) the fixed point took 100+ rounds instead of 3-8.
JPG_HD uint32_t span_bits_for(int nwords, int total_blocks, int max_threads) {
    const uint32_t nbits = (uint32_t)nwords * 32u;
    uint32_t s = (nbits + (uint32_t)max_threads - 1u) / (uint32_t)max_threads;
    const uint32_t per_block = nbits / (uint32_t)(total_blocks > 0 ? total_blocks : 1);
    if (s < 32u * per_block) s = 32u * per_block;
    s = (s + 31u) & ~31u;
    return s < 512u ? 512u : s;
}

#if !defined(__CUDACC__)
// Host statement of the parallel scheme (the kernel in esd_decode.cu runs the same decode_span with a thread per sub-sequence):
// returns the number of rounds the fixed point took.  Output as decode_scan_flat's.
// bstart != nullptr: the sparse hand-off (cf is then the entry list of `cap` + 1 entries, bstart has 6 n_mcus + 2 entries).
inline int decode_scan_parallel_host(const uint32_t* words, int nwords, const ScanTables& T, const int* td, const int* ta,
                                     const uint8_t* natural, int n_mcus, int16_t* cf, int max_threads, uint32_t* bstart = nullptr,
                                     int cap = 0) {
    const uint32_t S = span_bits_for(nwords, 6 * n_mcus, max_threads), nbits = (uint32_t)nwords * 32u;
    const int nt = (int)((nbits + S - 1) / S);
    const int total_blocks = 6 * n_mcus;
    SpanState* in = new SpanState[nt];
    SpanResult* res = new SpanResult[nt];
    auto limit = [&](int t) { const uint64_t e = (uint64_t)(t + 1) * S; return (uint32_t)(e < nbits ? e : nbits); };
    for (int t = 0; t < nt; ++t) {
        in[t] = SpanState{(uint32_t)t * S, 0, 0};
        res[t] = decode_span<SPAN_COUNT>(words, nwords, T, td, ta, natural, in[t], limit(t), nullptr, 0, nullptr, total_blocks);
    }
    int rounds = 1;
    for (;; ++rounds) {
        bool any = false;
        SpanState* nin = new SpanState[nt];
        for (int t = 0; t < nt; ++t) nin[t] = t == 0 ? in[0] : res[t - 1].end;   // Jacobi: everybody reads the previous round
        for (int t = 1; t < nt; ++t)
            if (!same_state(nin[t], in[t])) {
                in[t] = nin[t];
                res[t] = decode_span<SPAN_COUNT>(words, nwords, T, td, ta, natural, in[t], limit(t), nullptr, 0, nullptr, total_blocks);
                any = true;
            }
        delete[] nin;
        if (!any) break;
    }
    int blk = 0, pred[3] = {0, 0, 0}, o = 0;
    for (int t = 0; t < nt; ++t) {
        if (bstart) decode_span<SPAN_SPARSE>(words, nwords, T, td, ta, natural, in[t], limit(t), cf, blk, pred, total_blocks, bstart, o, cap);
        else decode_span<SPAN_DENSE>(words, nwords, T, td, ta, natural, in[t], limit(t), cf, blk, pred, total_blocks);
        blk += res[t].n_blocks;
        o += res[t].n_coefs;
        for (int c = 0; c < 3; ++c) pred[c] += res[t].dc[c];
    }
    delete[] in;
    delete[] res;
    return rounds;
}
#endif

// Where block `s` of the scan (decoding order) lies: component and block coordinates inside that component's plane.
JPG_HD void block_position(int s, int mcus_x, int* comp, int* bx, int* by) {
    const int mcu = s / 6, b = s - 6 * mcu;
    const int my = mcu / mcus_x, mx = mcu - my * mcus_x;
    *comp = b < 4 ? 0 : b - 3;
    *bx = b < 4 ? 2 * mx + (b & 1) : mx;
    *by = b < 4 ? 2 * my + (b >> 1) : my;
}

// Host helper: copies an entropy-coded segment up to its first marker with the byte stuffing removed.  Returns the number of
// bytes written; *clean is set to false when a restart marker was met (the flat decoder does not handle restarts).
inline size_t unstuff_scan(const uint8_t* src, size_t n, uint8_t* dst, bool* clean) {
    size_t o = 0, i = 0;
    *clean = true;
    while (i < n) {
        const uint8_t* ff = (const uint8_t*)memchr(src + i, 0xFF, n - i);
        const size_t run = ff ? (size_t)(ff - (src + i)) : n - i;
        memcpy(dst + o, src + i, run);
        o += run;
        i += run;
        if (!ff) break;
        if (i + 1 >= n) break;            // a lone FF at the end
        if (src[i + 1] == 0) { dst[o++] = 0xFF; i += 2; continue; }
        if (src[i + 1] == 0xFF) { ++i; continue; }  // fill byte
        if (src[i + 1] >= 0xD0 && src[i + 1] <= 0xD7) *clean = false;
        break;                            // a marker ends the segment
    }
    return o;
}

// ---- jidctint.c jpeg_idct_islow ------------------------------------------------------------------------------------
JPG_HD int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }
JPG_HD uint8_t range_limit(int32_t x) {  // sample + CENTERJSAMPLE clamped to 0..255
    x += 128;
    return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

// coef: 64 quantised coefficients (natural order); quant: 64 quantisation steps (natural order); out: 8 rows of 8 samples.
JPG_HD void idct_islow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int out_stride) {
    constexpr int CONST_BITS = 13, PASS1_BITS = 2;
    constexpr int32_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633, F_1_501 = 12299,
                      F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
    int32_t ws[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // pass 1: columns
        const int32_t i0 = coef[c] * (int32_t)quant[c], i1 = coef[8 + c] * (int32_t)quant[8 + c], i2 = coef[16 + c] * (int32_t)quant[16 + c],
                      i3 = coef[24 + c] * (int32_t)quant[24 + c], i4 = coef[32 + c] * (int32_t)quant[32 + c],
                      i5 = coef[40 + c] * (int32_t)quant[40 + c], i6 = coef[48 + c] * (int32_t)quant[48 + c],
                      i7 = coef[56 + c] * (int32_t)quant[56 + c];
        int32_t z2 = i2, z3 = i6;
        int32_t z1 = (z2 + z3) * F_0_541;
        int32_t tmp2 = z1 + z3 * (-F_1_847);
        int32_t tmp3 = z1 + z2 * F_0_765;
        z2 = i0; z3 = i4;
        int32_t tmp0 = (z2 + z3) << CONST_BITS;
        int32_t tmp1 = (z2 - z3) << CONST_BITS;
        const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = i7; tmp1 = i5; tmp2 = i3; tmp3 = i1;
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        const int32_t z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[c] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
        ws[56 + c] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
        ws[8 + c] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
        ws[48 + c] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
        ws[16 + c] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
        ws[40 + c] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
        ws[24 + c] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
        ws[32 + c] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // pass 2: rows
        const int32_t* w = ws + 8 * r;
        int32_t z2 = w[2], z3 = w[6];
        int32_t z1 = (z2 + z3) * F_0_541;
        int32_t tmp2 = z1 + z3 * (-F_1_847);
        int32_t tmp3 = z1 + z2 * F_0_765;
        int32_t tmp0 = (w[0] + w[4]) << CONST_BITS;
        int32_t tmp1 = (w[0] - w[4]) << CONST_BITS;
        const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        const int32_t z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        constexpr int SH = CONST_BITS + PASS1_BITS + 3;
        uint8_t* o = out + r * out_stride;
        o[0] = range_limit(descale(tmp10 + tmp3, SH));
        o[7] = range_limit(descale(tmp10 - tmp3, SH));
        o[1] = range_limit(descale(tmp11 + tmp2, SH));
        o[6] = range_limit(descale(tmp11 - tmp2, SH));
        o[2] = range_limit(descale(tmp12 + tmp1, SH));
        o[5] = range_limit(descale(tmp12 - tmp1, SH));
        o[3] = range_limit(descale(tmp13 + tmp0, SH));
        o[4] = range_limit(descale(tmp13 - tmp0, SH));
    }
}

// ---- jdsample.c h2v2_fancy_upsample + jdcolor.c ycc_rgb_convert -----------------------------------------------------------
// Chroma sample for output pixel (x, y) from the half-resolution plane `c` (cw x ch real samples, row stride cstride).
JPG_HD int colsum(const uint8_t* c, int cstride, int ch, int cy, int nb, int i) {
    (void)ch;
    return 3 * (int)c[cy * cstride + i] + (int)c[nb * cstride + i];
}
JPG_HD int fancy_chroma(const uint8_t* c, int cstride, int cw, int ch, int x, int y) {
    const int cy = y >> 1;
    int nb = (y & 1) ? cy + 1 : cy - 1;  // the nearer neighbouring chroma row; the picture's first / last row is replicated
    nb = nb < 0 ? 0 : (nb > ch - 1 ? ch - 1 : nb);
    const int i = x >> 1;
    const int cur = colsum(c, cstride, ch, cy, nb, i);
    if (!(x & 1)) {
        if (i == 0) return (cur * 4 + 8) >> 4;
        return (cur * 3 + colsum(c, cstride, ch, cy, nb, i - 1) + 8) >> 4;
    }
    if (i == cw - 1) return (cur * 4 + 7) >> 4;
    return (cur * 3 + colsum(c, cstride, ch, cy, nb, i + 1) + 7) >> 4;
}
JPG_HD uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
// JFIF YCbCr -> B, G, R exactly like jdcolor.c's tables (SCALEBITS 16)
JPG_HD void ycc_to_bgr(int y, int cb, int cr, uint8_t* bgr) {
    cb -= 128; cr -= 128;
    const int r = y + ((91881 * cr + 32768) >> 16);                   // FIX(1.40200)
    const int b = y + ((116130 * cb + 32768) >> 16);                  // FIX(1.77200)
    const int g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);     // FIX(0.34414), FIX(0.71414); ONE_HALF sits in Cb_g_tab
    bgr[0] = clamp255(b);
    bgr[1] = clamp255(g);
    bgr[2] = clamp255(r);
}

}  // namespace esdjpeg
