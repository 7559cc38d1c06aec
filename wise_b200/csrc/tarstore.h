// tarstore.h - host-side fast reader for WISE's WebdatasetStore shards.
// Replaces the per-sample python loop tarfile -> pickle.loads -> np.squeeze of
// /root/reference/src/feature/store/webdataset_store.py:116-141 (called from
// /root/reference/src/index/feature_search_index.py:79-82) with one pass over the mmap'ed shard:
// 512-byte ustar headers are walked directly and the float32 payload of each `%010d.features.pyd`
// member (pickle.dumps(np.ndarray (m, d) float32)) is located with a small pickle opcode scanner -
// no python objects are created per vector.  Pure host code (no CUDA), exported through the same C-ABI.
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace wbtar {

struct Sample {
    int64_t id;
    const unsigned char* data;  // m*d float32, little endian, C order
    int64_t m, d;
};

// Scan one numpy-array pickle (protocols 2-5): find the shape tuple and the raw data bytes object.
// Returns false for anything it does not fully understand (caller falls back to python).
inline bool scan_ndarray_pickle(const unsigned char* p, size_t n, Sample* out) {
    size_t i = 0;
    int64_t ints[8];
    int nints = 0;
    int64_t shape[2] = {0, 0};
    int ndim = -1;
    const unsigned char* best = nullptr;
    size_t best_len = 0;
    bool f4 = false, last_bool = false, fortran = false, order_f = false;
    auto need = [&](size_t k) { return i + k <= n; };
    while (i < n) {
        const unsigned char op = p[i++];
        switch (op) {
            case 0x80: if (!need(1)) return false; i += 1; nints = 0; break;                  // PROTO
            case 0x95: if (!need(8)) return false; i += 8; break;                             // FRAME
            case 0x8c: { if (!need(1)) return false; size_t l = p[i++]; if (!need(l)) return false;  // SHORT_BINUNICODE
                         if (l == 2 && p[i] == 'f' && p[i + 1] == '4') f4 = true;
                         if (l == 1 && p[i] == 'F') order_f = true;  // _frombuffer(..., order='F') (proto 5)
                         i += l; nints = 0; break; }
            case 'X': { if (!need(4)) return false; uint32_t l; memcpy(&l, p + i, 4); i += 4; if (!need(l)) return false;
                        if (l == 2 && p[i] == 'f' && p[i + 1] == '4') f4 = true;
                        i += l; nints = 0; break; }
            case 0x8d: { if (!need(8)) return false; uint64_t l; memcpy(&l, p + i, 8); i += 8; if (!need(l)) return false; i += l; nints = 0; break; }
            case 'U': { if (!need(1)) return false; size_t l = p[i++]; if (!need(l)) return false;   // SHORT_BINSTRING (proto 2)
                        if (l == 2 && p[i] == 'f' && p[i + 1] == '4') f4 = true;
                        i += l; nints = 0; break; }
            case 'c': { // GLOBAL: two newline-terminated lines
                int nl = 0; while (i < n && nl < 2) { if (p[i++] == '\n') nl++; } nints = 0; break; }
            case 0x94: break;                                                                  // MEMOIZE
            case 0x93: case 'R': case 'N': case ')': case '(': case 0x81: nints = 0; break;    // STACK_GLOBAL REDUCE NONE EMPTY_TUPLE MARK NEWOBJ
            case 0x88: last_bool = true; nints = 0; break;                                     // NEWTRUE
            case 0x89: last_bool = false; nints = 0; break;                                    // NEWFALSE
            case 'K': if (!need(1)) return false; if (nints < 8) ints[nints++] = p[i]; i += 1; break;                          // BININT1
            case 'M': { if (!need(2)) return false; uint16_t v; memcpy(&v, p + i, 2); i += 2; if (nints < 8) ints[nints++] = v; break; }  // BININT2
            case 'J': { if (!need(4)) return false; int32_t v; memcpy(&v, p + i, 4); i += 4; if (nints < 8) ints[nints++] = v; break; }   // BININT
            case 0x8a: { if (!need(1)) return false; size_t l = p[i++]; if (!need(l) || l > 8) return false;                    // LONG1
                         int64_t v = 0; memcpy(&v, p + i, l); i += l; if (nints < 8) ints[nints++] = v; break; }
            case 0x85: case 0x86: case 0x87: {                                                 // TUPLE1/2/3
                const int k = op - 0x84;
                if (nints >= k && k <= 2) {  // the last all-int tuple before the data bytes is the shape
                    if (k == 1) { shape[0] = ints[nints - 1]; shape[1] = 1; ndim = 1; }
                    else { shape[0] = ints[nints - 2]; shape[1] = ints[nints - 1]; ndim = 2; }
                }
                nints = 0; break; }
            case 't': nints = 0; break;                                                        // TUPLE (from MARK)
            case 'h': if (!need(1)) return false; i += 1; nints = 0; break;                    // BINGET
            case 'j': if (!need(4)) return false; i += 4; nints = 0; break;                    // LONG_BINGET
            case 'q': if (!need(1)) return false; i += 1; break;                               // BINPUT
            case 'r': if (!need(4)) return false; i += 4; break;                               // LONG_BINPUT
            case 'C': { if (!need(1)) return false; size_t l = p[i++]; if (!need(l)) return false;
                        if (l >= best_len) { best = p + i; best_len = l; fortran = last_bool; } i += l; nints = 0; break; }
            case 'B': { if (!need(4)) return false; uint32_t l; memcpy(&l, p + i, 4); i += 4; if (!need(l)) return false;
                        if (l >= best_len) { best = p + i; best_len = l; fortran = last_bool; } i += l; nints = 0; break; }
            case 0x96: { if (!need(8)) return false; uint64_t l; memcpy(&l, p + i, 8); i += 8; if (!need(l)) return false;  // BYTEARRAY8 (proto 5)
                         if (l >= best_len) { best = p + i; best_len = (size_t)l; fortran = last_bool; } i += l; nints = 0; break; }
            case 0x8e: { if (!need(8)) return false; uint64_t l; memcpy(&l, p + i, 8); i += 8; if (!need(l)) return false;
                         if (l >= best_len) { best = p + i; best_len = (size_t)l; fortran = last_bool; } i += l; nints = 0; break; }
            case 'T': { if (!need(4)) return false; uint32_t l; memcpy(&l, p + i, 4); i += 4; if (!need(l)) return false;  // BINSTRING (proto 2 data)
                        if (l >= best_len) { best = p + i; best_len = l; fortran = last_bool; } i += l; nints = 0; break; }
            case 'b': nints = 0; break;                                                        // BUILD
            case '.': i = n; break;                                                            // STOP
            default: return false;  // unknown opcode: let python handle this sample
        }
    }
    // the boolean right before the data bytes is the state tuple's is_fortran flag
    if (!best || !f4 || fortran || order_f || ndim < 1) return false;
    if (ndim == 1) { shape[1] = shape[0]; shape[0] = 1; }  // a bare (d,) vector counts as one row
    if (shape[0] <= 0 || shape[1] <= 0 || (size_t)(shape[0] * shape[1] * 4) != best_len) return false;
    out->data = best;
    out->m = shape[0];
    out->d = shape[1];
    return true;
}

struct Mapped {
    const unsigned char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const unsigned char*)m;
        return true;
    }
    ~Mapped() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

// Walk the shard; call fn(sample) for every `<int>.features.pyd` member. Returns 0 ok, 1 io error,
// 2 a member this reader does not understand (python fallback), 3 malformed tar.
template <class F>
inline int walk(const char* path, F fn, std::string* why) {
    Mapped mp;
    if (!mp.open(path)) { *why = std::string("cannot open ") + path; return 1; }
    size_t off = 0;
    std::string longname;
    while (off + 512 <= mp.n) {
        const unsigned char* h = mp.p + off;
        bool zero = true;
        for (int i = 0; i < 512 && zero; ++i) zero = h[i] == 0;
        if (zero) break;  // end-of-archive marker
        char name[257];
        size_t nl = strnlen((const char*)h, 100);
        memcpy(name, h, nl);
        name[nl] = 0;
        uint64_t size = 0;
        if (h[124] & 0x80) {  // GNU base-256 size
            for (int i = 125; i < 136; ++i) size = (size << 8) | h[i];
        } else {
            for (int i = 124; i < 136 && h[i] >= '0' && h[i] <= '7'; ++i) size = size * 8 + (h[i] - '0');
        }
        const char type = (char)h[156];
        const size_t data = off + 512;
        if (data + size > mp.n) { *why = "truncated tar member"; return 3; }
        const size_t next = data + ((size + 511) & ~(size_t)511);
        if (type == 'L') {  // GNU long name for the next member
            longname.assign((const char*)mp.p + data, strnlen((const char*)mp.p + data, size));
        } else if (type == '0' || type == 0) {
            std::string nm = longname.empty() ? std::string(name) : longname;
            longname.clear();
            const size_t slash = nm.find_last_of('/');
            const std::string base = slash == std::string::npos ? nm : nm.substr(slash + 1);
            const size_t dot = base.find('.');
            if (dot != std::string::npos && base.substr(dot + 1) == "features.pyd") {
                Sample s{};
                char* endp = nullptr;
                s.id = strtoll(base.c_str(), &endp, 10);
                if (endp != base.c_str() + dot) { *why = "non-integer sample key " + base; return 2; }
                if (!scan_ndarray_pickle(mp.p + data, size, &s)) { *why = "unsupported pickle in " + base; return 2; }
                const int rc = fn(s);
                if (rc) return rc;
            }
        } else {
            longname.clear();
        }
        off = next;
    }
    return 0;
}

}  // namespace wbtar
