// Host-side JPEG decoder behind the compat layer's load_image_color (darknet/src/image.c:1442-1482 decodes through
// stb_image; models_detection/YOLO.py:141 always hands a .jpg path to it).
//
// Scope: Huffman-coded DCT JPEG, 8-bit: sequential (SOF0 baseline / SOF1 extended) and progressive (SOF2: spectral
// selection + successive approximation, DC/AC first and refinement scans, EOB runs), grey or 3 components, sampling
// factors 1..4, interleaved and non-interleaved scans, restart intervals, JFIF / Adobe colour-transform markers.
// Arithmetic-coded, lossless and 12-bit files are rejected with an error (decode them in the caller and use
// make_image()).
//
// The reconstruction is bit-exact with the reference's decoder, which matters because the frame feeds a detector
// whose parity is checked against the reference library:
//   * IDCT: the 13-bit fixed-point "islow" factorisation (IJG jidctint), column pass keeps 2 extra bits
//     ((x + 512) >> 10), row pass rounds and re-centres in one shift ((x + 65536 + (128 << 17)) >> 17), clamp to 0..255;
//   * chroma upsampling: triangle filters -- (3*near + far + 2) >> 2 along one axis, (3*t0 + t1 + 8) >> 4 with
//     t = 3*near + far for 2x2 -- and nearest-neighbour for every other factor; the row pairing (which chroma row is
//     "near") follows the half-sample-centred phase of the reference;
//   * YCbCr -> RGB: 20-bit fixed point with coefficients rounded to 12 bits then shifted by 8, the green Cb term masked
//     to its upper 16 bits, y biased by 1 << 19.
// tests/test_jpeg_cpu.py pins all of it against oracle/_ref/libdarknet.so (the reference's own stb path) and against
// committed outputs of that library.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace b2t {

namespace {

struct Huff {
    // canonical code tables (ITU T.81 Annex C / F.2.2.3): codes of length l are [mincode[l], maxcode[l]], their symbols
    // start at valptr[l]
    uint8_t symbols[256];
    int mincode[17], maxcode[18], valptr[17];
    bool defined = false;
    uint16_t fast[512];          // 9-bit prefix -> (length << 8) | symbol, 0 = longer code
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int w = 0, hgt = 0;          // real size in samples
    int w2 = 0, h2 = 0;          // allocated size (whole MCUs)
    int dc_pred = 0;
    std::vector<uint8_t> data;
    std::vector<int16_t> coeff;  // progressive: every block's 64 coefficients (natural order), bw2 blocks per row
    int bw2 = 0;
};

struct Decoder {
    const uint8_t *p, *end;
    std::string error;
    int width = 0, height = 0, ncomp = 0;
    int hmax = 1, vmax = 1, mcu_w = 8, mcu_h = 8, mcux = 0, mcuy = 0;
    uint16_t qt[4][64];
    Huff dc[4], ac[4];
    Component comp[4];
    int restart_interval = 0;
    bool jfif = false;
    int adobe_transform = -1;
    bool progressive = false;
    int ss = 0, se = 63, ah = 0, al = 0, eob_run = 0;     // progressive scan parameters
    // bit reader
    uint32_t bits = 0;
    int nbits = 0;
    bool hit_marker = false;
    uint8_t marker = 0;

    bool fail(const char *msg) { if (error.empty()) error = msg; return false; }
    int u8() { return p < end ? *p++ : 0; }
    int u16() { const int a = u8(); return (a << 8) | u8(); }
};

const uint8_t kZigzag[64 + 16] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};   // overrun guard

bool build_huffman(Huff &h, const uint8_t counts[16], const uint8_t *symbols, int n) {
    memcpy(h.symbols, symbols, n);
    memset(h.fast, 0, sizeof h.fast);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h.valptr[l] = k;
        h.mincode[l] = code;
        for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code)
            if (l <= 9) {
                const int first = code << (9 - l);
                for (int j = 0; j < (1 << (9 - l)); ++j) h.fast[first + j] = (uint16_t)((l << 8) | symbols[k]);
            }
        h.maxcode[l] = counts[l - 1] ? code - 1 : -1;
        if (code > (1 << l)) return false;
        code <<= 1;
    }
    h.maxcode[17] = 0x7fffffff;
    h.defined = true;
    return true;
}

// ---- entropy-coded segment: bytes with 0xFF00 stuffing; a marker ends the data (zeros are fed from there on)
void fill(Decoder &d) {
    while (d.nbits <= 24) {
        int b = 0;
        if (!d.hit_marker && d.p < d.end) {
            b = *d.p++;
            if (b == 0xFF) {
                int m = d.p < d.end ? *d.p++ : 0xD9;
                while (m == 0xFF && d.p < d.end) m = *d.p++;           // fill bytes
                if (m != 0) { d.marker = (uint8_t)m; d.hit_marker = true; b = 0; }
            }
        }
        d.bits |= (uint32_t)b << (24 - d.nbits);
        d.nbits += 8;
    }
}
inline int peek(Decoder &d, int n) { return (int)(d.bits >> (32 - n)); }
inline void skip(Decoder &d, int n) { d.bits <<= n; d.nbits -= n; }

int decode_symbol(Decoder &d, const Huff &h) {
    if (d.nbits < 16) fill(d);
    const uint16_t f = h.fast[peek(d, 9)];
    if (f) { skip(d, f >> 8); return f & 255; }
    int code = peek(d, 10), l = 10;
    for (; l <= 16; ++l) {
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) break;
        code = peek(d, l + 1);
    }
    if (l > 16) return -1;
    skip(d, l);
    return h.symbols[h.valptr[l] + code - h.mincode[l]];
}

// n additional bits -> signed value (T.81 F.2.2.1 EXTEND)
inline int receive_extend(Decoder &d, int n) {
    if (n == 0) return 0;
    if (d.nbits < n) fill(d);
    const int v = peek(d, n);
    skip(d, n);
    return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
}

bool decode_block(Decoder &d, Component &c, int16_t blk[64]) {
    memset(blk, 0, 64 * sizeof(int16_t));
    const uint16_t *q = d.qt[c.tq];
    int t = decode_symbol(d, d.dc[c.td]);
    if (t < 0 || t > 15) return d.fail("bad huffman code");
    c.dc_pred += receive_extend(d, t);
    blk[0] = (int16_t)(c.dc_pred * q[0]);
    for (int k = 1; k < 64;) {
        const int rs = decode_symbol(d, d.ac[c.ta]);
        if (rs < 0) return d.fail("bad huffman code");
        const int r = rs >> 4, s = rs & 15;
        if (s == 0) {
            if (rs != 0xF0) break;          // end of block
            k += 16;
        } else {
            k += r;
            const int z = kZigzag[k++];
            blk[z] = (int16_t)(receive_extend(d, s) * q[z]);
        }
    }
    return true;
}

// ---- progressive scans (T.81 annex G): coefficients are accumulated in Component::coeff over several scans
inline int get_bit(Decoder &d) {
    if (d.nbits < 1) fill(d);
    const int v = peek(d, 1);
    skip(d, 1);
    return v;
}
inline int get_bits(Decoder &d, int n) {
    if (n == 0) return 0;
    if (d.nbits < n) fill(d);
    const int v = peek(d, n);
    skip(d, n);
    return v;
}

bool decode_block_prog_dc(Decoder &d, Component &c, int16_t *blk) {
    if (d.se != 0) return d.fail("progressive DC scan with an AC band");
    if (d.ah == 0) {                                   // first pass: the (point-transformed) DC difference
        memset(blk, 0, 64 * sizeof(int16_t));
        const int t = decode_symbol(d, d.dc[c.td]);
        if (t < 0 || t > 15) return d.fail("bad huffman code");
        c.dc_pred += receive_extend(d, t);
        blk[0] = (int16_t)(c.dc_pred * (1 << d.al));
    } else if (get_bit(d)) {                           // refinement: one more bit of precision
        blk[0] = (int16_t)(blk[0] + (1 << d.al));
    }
    return true;
}

bool decode_block_prog_ac(Decoder &d, Component &c, int16_t *blk) {
    if (d.ss == 0) return d.fail("progressive AC scan starting at the DC coefficient");
    const Huff &h = d.ac[c.ta];
    if (d.ah == 0) {                                   // first pass of the band [ss, se]
        if (d.eob_run) { --d.eob_run; return true; }
        int k = d.ss;
        do {
            const int rs = decode_symbol(d, h);
            if (rs < 0) return d.fail("bad huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r < 15) {                          // end-of-band run of 2^r + extra blocks (this one included)
                    d.eob_run = 1 << r;
                    if (r) d.eob_run += get_bits(d, r);
                    --d.eob_run;
                    break;
                }
                k += 16;
            } else {
                k += r;
                const int z = kZigzag[k++];
                blk[z] = (int16_t)(receive_extend(d, s) * (1 << d.al));
            }
        } while (k <= d.se);
        return true;
    }
    // refinement pass: correction bits for the coefficients that are already non-zero, new +-1 coefficients between them
    const int16_t bit = (int16_t)(1 << d.al);
    auto refine = [&](int16_t *p) {
        if (get_bit(d) && (*p & bit) == 0) *p = (int16_t)(*p > 0 ? *p + bit : *p - bit);
    };
    if (d.eob_run) {
        --d.eob_run;
        for (int k = d.ss; k <= d.se; ++k) {
            int16_t *p = &blk[kZigzag[k]];
            if (*p != 0) refine(p);
        }
        return true;
    }
    int k = d.ss;
    do {
        const int rs = decode_symbol(d, h);
        if (rs < 0) return d.fail("bad huffman code");
        int s = rs & 15, r = rs >> 4;
        if (s == 0) {
            if (r < 15) {
                d.eob_run = (1 << r) - 1;
                if (r) d.eob_run += get_bits(d, r);
                r = 64;                                // run to the end of the band, refining on the way
            }                                          // r == 15: sixteen zero coefficients to skip
        } else {
            if (s != 1) return d.fail("bad huffman code");
            s = get_bit(d) ? bit : -bit;
        }
        while (k <= d.se) {
            int16_t *p = &blk[kZigzag[k++]];
            if (*p != 0) {
                refine(p);
            } else {
                if (r == 0) { *p = (int16_t)s; break; }
                --r;
            }
        }
    } while (k <= d.se);
    return true;
}

// ---- inverse DCT (see the header comment).  One 8-point pass on already scaled inputs; outputs the even part x[0..3]
// and the odd part t[0..3] such that out[k] = x[k] + t[3-k], out[7-k] = x[k] - t[3-k].
// The multipliers are the usual islow constants scaled by 2^12; each is the value (int)(c * 4096 + 0.5) of the SIGNED
// constant as the reference evaluates it (conversion to int truncates toward zero, so the negative ones are one closer
// to zero than -round(|c| * 4096)):
//   0.5411961 -> 2217   -1.847759065 -> -7567   0.765366865 -> 3135   1.175875602 -> 4816   0.298631336 -> 1223
//   2.053119869 -> 8410   3.072711026 -> 12586   1.501321110 -> 6149   -0.899976223 -> -3685   -2.562915447 -> -10497
//   -1.961570560 -> -8034   -0.390180644 -> -1597
inline void idct8(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int x[4], int t[4]) {
    const int z = (s2 + s6) * 2217;
    const int e2 = z + s6 * -7567, e3 = z + s2 * 3135;
    const int e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
    x[0] = e0 + e3; x[3] = e0 - e3; x[1] = e1 + e2; x[2] = e1 - e2;
    const int a = s7 + s3, b = s5 + s1, c = s7 + s1, dd = s5 + s3;
    const int z5 = (a + b) * 4816;
    const int c1 = z5 + c * -3685, c2 = z5 + dd * -10497;
    const int a2 = a * -8034, b2 = b * -1597;
    t[3] = s1 * 6149 + c1 + b2;
    t[2] = s3 * 12586 + c2 + a2;
    t[1] = s5 * 8410 + c2 + b2;
    t[0] = s7 * 1223 + c1 + a2;
}
inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

void idct_block(uint8_t *out, int stride, const int16_t d[64]) {
    int v[64], x[4], t[4];
    for (int i = 0; i < 8; ++i) {                       // columns, 2 extra bits kept
        idct8(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i], x, t);
        for (int k = 0; k < 4; ++k) {
            v[8 * k + i] = (x[k] + 512 + t[3 - k]) >> 10;
            v[8 * (7 - k) + i] = (x[k] + 512 - t[3 - k]) >> 10;
        }
    }
    for (int i = 0; i < 8; ++i, out += stride) {        // rows: remove 2^17, round, re-centre by +128
        const int *r = v + 8 * i;
        idct8(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], x, t);
        for (int k = 0; k < 4; ++k) {
            const int bias = x[k] + 65536 + (128 << 17);
            out[k] = clamp8((bias + t[3 - k]) >> 17);
            out[7 - k] = clamp8((bias - t[3 - k]) >> 17);
        }
    }
}

// ---- markers
bool read_dqt(Decoder &d, int len) {
    while (len > 0) {
        const int pq = d.u8(), prec = pq >> 4, id = pq & 15;
        if (id > 3 || prec > 1) return d.fail("bad DQT");
        for (int i = 0; i < 64; ++i) d.qt[id][kZigzag[i]] = (uint16_t)(prec ? d.u16() : d.u8());
        len -= 65 + 64 * prec;
    }
    return len == 0 || d.fail("bad DQT length");
}

bool read_dht(Decoder &d, int len) {
    while (len > 0) {
        const int tc = d.u8(), cls = tc >> 4, id = tc & 15;
        if (cls > 1 || id > 3) return d.fail("bad DHT");
        uint8_t counts[16], symbols[256];
        int n = 0;
        for (int i = 0; i < 16; ++i) { counts[i] = (uint8_t)d.u8(); n += counts[i]; }
        if (n > 256) return d.fail("bad DHT");
        for (int i = 0; i < n; ++i) symbols[i] = (uint8_t)d.u8();
        if (!build_huffman(cls ? d.ac[id] : d.dc[id], counts, symbols, n)) return d.fail("bad huffman code lengths");
        len -= 17 + n;
    }
    return len == 0 || d.fail("bad DHT length");
}

bool read_sof(Decoder &d) {
    d.u16();
    if (d.u8() != 8) return d.fail("only 8-bit JPEG is supported");
    d.height = d.u16(); d.width = d.u16(); d.ncomp = d.u8();
    if (d.width < 1 || d.height < 1) return d.fail("bad JPEG size");
    if (d.ncomp != 1 && d.ncomp != 3) return d.fail("only grey and 3-component JPEG are supported");
    for (int i = 0; i < d.ncomp; ++i) {
        Component &c = d.comp[i];
        c.id = d.u8();
        const int hv = d.u8();
        c.h = hv >> 4; c.v = hv & 15; c.tq = d.u8();
        if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return d.fail("bad component");
        if (c.h > d.hmax) d.hmax = c.h;
        if (c.v > d.vmax) d.vmax = c.v;
    }
    for (int i = 0; i < d.ncomp; ++i)
        if (d.hmax % d.comp[i].h || d.vmax % d.comp[i].v) return d.fail("fractional sampling ratios are not supported");
    d.mcu_w = 8 * d.hmax; d.mcu_h = 8 * d.vmax;
    d.mcux = (d.width + d.mcu_w - 1) / d.mcu_w; d.mcuy = (d.height + d.mcu_h - 1) / d.mcu_h;
    for (int i = 0; i < d.ncomp; ++i) {
        Component &c = d.comp[i];
        c.w = (d.width * c.h + d.hmax - 1) / d.hmax; c.hgt = (d.height * c.v + d.vmax - 1) / d.vmax;
        c.w2 = d.mcux * c.h * 8; c.h2 = d.mcuy * c.v * 8;
        c.data.assign((size_t)c.w2 * c.h2, 0);
        if (d.progressive) {
            c.bw2 = c.w2 / 8;
            c.coeff.assign((size_t)c.w2 * c.h2, 0);
        }
    }
    return true;
}

void restart(Decoder &d) {
    d.bits = 0; d.nbits = 0; d.hit_marker = false; d.marker = 0; d.eob_run = 0;
    for (int i = 0; i < 4; ++i) d.comp[i].dc_pred = 0;
}

bool read_scan(Decoder &d) {
    d.u16();
    const int ns = d.u8();
    if (ns < 1 || ns > d.ncomp) return d.fail("bad SOS");
    int order[4];
    for (int i = 0; i < ns; ++i) {
        const int id = d.u8(), tt = d.u8();
        int k = 0;
        while (k < d.ncomp && d.comp[k].id != id) ++k;
        if (k == d.ncomp) return d.fail("bad SOS component");
        d.comp[k].td = tt >> 4; d.comp[k].ta = tt & 15;
        if (d.comp[k].td > 3 || d.comp[k].ta > 3) return d.fail("bad huffman table index");
        order[i] = k;
    }
    d.ss = d.u8(); d.se = d.u8();
    const int a = d.u8();
    d.ah = a >> 4; d.al = a & 15;
    if (d.progressive) {
        if (d.ss > 63 || d.se > 63 || d.ss > d.se || d.ah > 13 || d.al > 13) return d.fail("bad progressive scan parameters");
        if (d.ss != 0 && ns != 1) return d.fail("interleaved progressive AC scan");
    } else {
        d.ss = 0; d.se = 63; d.ah = d.al = 0;
    }
    for (int i = 0; i < ns; ++i) {
        const Component &c = d.comp[order[i]];
        const bool need_dc = d.ss == 0, need_ac = d.se != 0;
        if ((need_dc && !(d.progressive && d.ah) && !d.dc[c.td].defined) || (need_ac && !d.ac[c.ta].defined))
            return d.fail("scan uses an undefined huffman table");
    }
    restart(d);
    int16_t blk[64];
    int todo = d.restart_interval ? d.restart_interval : 0x7fffffff;
    auto after_unit = [&]() -> bool {
        if (--todo > 0) return true;
        // restart interval exhausted: the next thing in the stream must be RSTn
        if (d.nbits < 24) fill(d);
        if (!(d.hit_marker && d.marker >= 0xD0 && d.marker <= 0xD7)) return false;   // no restart marker: scan ends here
        restart(d);
        todo = d.restart_interval;
        return true;
    };
    if (ns == 1) {                   // non-interleaved: the component's own blocks in raster order
        Component &c = d.comp[order[0]];
        const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3;
        for (int by = 0; by < bh; ++by)
            for (int bx = 0; bx < bw; ++bx) {
                if (d.progressive) {
                    int16_t *cb = c.coeff.data() + 64 * ((size_t)by * c.bw2 + bx);
                    if (!(d.ss == 0 ? decode_block_prog_dc(d, c, cb) : decode_block_prog_ac(d, c, cb))) return false;
                } else {
                    if (!decode_block(d, c, blk)) return false;
                    idct_block(c.data.data() + (size_t)by * 8 * c.w2 + bx * 8, c.w2, blk);
                }
                if (!after_unit()) return true;
            }
        return true;
    }
    for (int my = 0; my < d.mcuy; ++my)
        for (int mx = 0; mx < d.mcux; ++mx) {
            for (int i = 0; i < ns; ++i) {
                Component &c = d.comp[order[i]];
                for (int y = 0; y < c.v; ++y)
                    for (int x = 0; x < c.h; ++x) {
                        if (d.progressive) {           // only DC scans may be interleaved
                            int16_t *cb = c.coeff.data() + 64 * ((size_t)(my * c.v + y) * c.bw2 + (mx * c.h + x));
                            if (!decode_block_prog_dc(d, c, cb)) return false;
                            continue;
                        }
                        if (!decode_block(d, c, blk)) return false;
                        idct_block(c.data.data() + (size_t)(my * c.v + y) * 8 * c.w2 + (mx * c.h + x) * 8, c.w2, blk);
                    }
            }
            if (!after_unit()) return true;
        }
    return true;
}

// ---- upsampling of one output row: returns the row to use (may be `near` itself)
const uint8_t *upsample_row(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) return near;
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        return out;
    }
    if (hs == 2 && vs == 1) {
        if (w == 1) { out[0] = out[1] = near[0]; return out; }
        out[0] = near[0];
        out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
        for (int i = 1; i < w - 1; ++i) {
            out[2 * i] = (uint8_t)((3 * near[i] + near[i - 1] + 2) >> 2);
            out[2 * i + 1] = (uint8_t)((3 * near[i] + near[i + 1] + 2) >> 2);
        }
        out[2 * w - 2] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
        out[2 * w - 1] = near[w - 1];
        return out;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) { out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2); return out; }
        int t1 = 3 * near[0] + far[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
    return out;
}


}  // namespace

// -> interleaved RGB, 3 bytes per pixel.  Returns false and fills `error` on failure.
bool jpeg_decode_rgb(const uint8_t *data, size_t size, std::vector<uint8_t> &rgb, int &width, int &height, std::string &error) {
    Decoder d;
    d.p = data; d.end = data + size;
    memset(d.qt, 0, sizeof d.qt);
    auto bad = [&](const char *m) { error = std::string("jpeg: ") + (d.error.empty() ? m : d.error.c_str()); return false; };
    if (size < 4 || d.u8() != 0xFF || d.u8() != 0xD8) return bad("not a JPEG file");
    bool have_frame = false, have_scan = false, is_rgb_ids = false;
    for (;;) {
        int m;
        if (d.hit_marker) { m = d.marker; d.hit_marker = false; d.marker = 0; }
        else {
            int b = d.u8();
            while (b != 0xFF && d.p < d.end) b = d.u8();
            if (d.p >= d.end) break;
            m = d.u8();
            while (m == 0xFF && d.p < d.end) m = d.u8();
        }
        if (m == 0xD9) break;                                   // EOI
        if (m == 0 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (have_frame) return bad("second frame header");
            d.progressive = m == 0xC2;
            if (!read_sof(d)) return bad("bad frame header");
            have_frame = true;
            is_rgb_ids = d.ncomp == 3 && d.comp[0].id == 'R' && d.comp[1].id == 'G' && d.comp[2].id == 'B';
            continue;
        }
        if ((m >= 0xC3 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) return bad("unsupported JPEG coding process");
        if (m == 0xDA) {
            if (!have_frame) return bad("scan before frame header");
            d.nbits = 0; d.bits = 0;
            if (!read_scan(d)) return bad("corrupt scan");
            have_scan = true;
            continue;
        }
        const int len = d.u16() - 2;
        if (len < 0 || d.p + len > d.end) return bad("truncated marker segment");
        const uint8_t *seg_end = d.p + len;
        if (m == 0xDB) { if (!read_dqt(d, len)) return bad("bad DQT"); }
        else if (m == 0xC4) { if (!read_dht(d, len)) return bad("bad DHT"); }
        else if (m == 0xDD) d.restart_interval = d.u16();
        else if (m == 0xE0 && len >= 5 && !memcmp(d.p, "JFIF\0", 5)) d.jfif = true;
        else if (m == 0xEE && len >= 12 && !memcmp(d.p, "Adobe\0", 6)) d.adobe_transform = d.p[11];
        d.p = seg_end;
    }
    if (!have_frame || !have_scan) return bad("no image data");
    if (d.progressive)                                  // all scans are in: dequantise and transform every block
        for (int k = 0; k < d.ncomp; ++k) {
            Component &c = d.comp[k];
            const uint16_t *q = d.qt[c.tq];
            const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3;
            for (int by = 0; by < bh; ++by)
                for (int bx = 0; bx < bw; ++bx) {
                    int16_t *cb = c.coeff.data() + 64 * ((size_t)by * c.bw2 + bx);
                    for (int i = 0; i < 64; ++i) cb[i] = (int16_t)(cb[i] * q[i]);
                    idct_block(c.data.data() + (size_t)by * 8 * c.w2 + bx * 8, c.w2, cb);
                }
        }

    // resample + colour conversion, row by row
    width = d.width; height = d.height;
    rgb.assign((size_t)width * height * 3, 0);
    const bool is_rgb = d.ncomp == 3 && (is_rgb_ids || (d.adobe_transform == 0 && !d.jfif));
    struct Res { int hs, vs, ystep, ypos, w_lo; const uint8_t *line0, *line1; std::vector<uint8_t> buf; } res[3];
    for (int k = 0; k < d.ncomp; ++k) {
        Res &r = res[k];
        r.hs = d.hmax / d.comp[k].h; r.vs = d.vmax / d.comp[k].v;
        r.ystep = r.vs >> 1; r.ypos = 0;
        r.w_lo = (width + r.hs - 1) / r.hs;
        r.line0 = r.line1 = d.comp[k].data.data();
        r.buf.assign((size_t)width + 8, 0);
    }
    // ((int)(c * 4096 + 0.5)) << 8 for c = 1.402, 0.71414, 0.34414, 1.772
    const int kr = 1470208, kg_cr = 748800, kg_cb = 360960, kb = 1858048;
    for (int j = 0; j < height; ++j) {
        const uint8_t *row[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < d.ncomp; ++k) {
            Res &r = res[k];
            const bool bottom = r.ystep >= (r.vs >> 1);         // which of the two source rows is the nearer one
            row[k] = upsample_row(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lo, r.hs, r.vs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < d.comp[k].hgt) r.line1 += d.comp[k].w2;
            }
        }
        uint8_t *o = rgb.data() + (size_t)j * width * 3;
        if (d.ncomp == 1) {
            for (int i = 0; i < width; ++i, o += 3) o[0] = o[1] = o[2] = row[0][i];
        } else if (is_rgb) {
            for (int i = 0; i < width; ++i, o += 3) { o[0] = row[0][i]; o[1] = row[1][i]; o[2] = row[2][i]; }
        } else {
            for (int i = 0; i < width; ++i, o += 3) {
                const int yf = (row[0][i] << 20) + (1 << 19);
                const int cb = row[1][i] - 128, cr = row[2][i] - 128;
                int r = yf + cr * kr;
                int g = yf + cr * -kg_cr + (int)((uint32_t)(cb * -kg_cb) & 0xffff0000u);
                int b = yf + cb * kb;
                r >>= 20; g >>= 20; b >>= 20;
                o[0] = clamp8(r); o[1] = clamp8(g); o[2] = clamp8(b);
            }
        }
    }
    return true;
}

}  // namespace b2t
