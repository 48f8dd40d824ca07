// Minimal native HDF5 reader for CLAM-style bag files (host code; no libhdf5, no h5py).
//
// The reference reads  h5_files/<slide>.h5['features'] [N,512] float32  and ['coords'] [N,2]  through h5py on every
// access (datasets/dataset_generic.py:424-430); the files are written by CLAM's save_hdf5 (utils/file_utils.py:16-35):
// h5py defaults (libver "earliest": superblock v0, version-1 object headers, symbol-table groups), chunked layout
// with chunk shape (1, ...), resizable along axis 0, no compression.  This reader implements exactly that subset of
// the HDF5 file format (and contiguous / compact layouts, superblock v1, layout message versions 1-3):
//   superblock -> root symbol-table entry -> group B-tree (v1, type 0) + local heap + SNOD nodes -> object header
//   (v1, with continuation blocks) -> dataspace / datatype / layout / filter messages -> chunk B-tree (v1, type 1).
// Anything else (superblock v2/v3, "OHDR" object headers, filters, variable-length types) is reported as
// unsupported rather than guessed at.  Files are mmap-ed; every address is bounds-checked against the mapping.
// Host pointers only: this file launches no kernels.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace moc {
namespace {

constexpr uint64_t H5_UNDEF = ~0ull;

struct H5File {
    const uint8_t* map = nullptr;
    size_t size = 0;
    uint64_t base = 0;       // file address of the superblock's base address (user block size)
    int so = 8, sl = 8;      // size of offsets / lengths
    uint64_t root_btree = H5_UNDEF, root_heap = H5_UNDEF, root_ohdr = H5_UNDEF;
};

struct H5Dataset {
    int rank = 0;
    uint64_t dims[4] = {0, 0, 0, 0};
    int type_class = -1, elem_size = 0, big_endian = 0;
    int layout = -1;  // 0 compact, 1 contiguous, 2 chunked
    uint64_t data_addr = H5_UNDEF, data_size = 0;
    const uint8_t* compact = nullptr;
    int chunk_rank = 0;
    uint32_t chunk_dims[5] = {0, 0, 0, 0, 0};
    uint64_t chunk_btree = H5_UNDEF;
    bool filtered = false;
};

struct Reader {
    const H5File& f;
    bool ok = true;
    explicit Reader(const H5File& file) : f(file) {}
    const uint8_t* at(uint64_t addr, uint64_t len) {  // addr relative to the base address
        if (addr == H5_UNDEF || f.base + addr + len > f.size || f.base + addr + len < f.base + addr) {
            ok = false;
            return nullptr;
        }
        return f.map + f.base + addr;
    }
    static uint64_t le(const uint8_t* p, int n) {
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
        if (n < 8 && v == ((1ull << (8 * n)) - 1)) return H5_UNDEF;  // all-ones = undefined address
        return v;
    }
    static uint64_t le_raw(const uint8_t* p, int n) {
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
        return v;
    }
};

int fail(const char* what) {
    set_error("moc_h5: %s", what);
    return MOC_E_ARG;
}

int parse_superblock(H5File& f) {
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    uint64_t off = 0;
    bool found = false;
    for (; off + 8 <= f.size; off = off ? off * 2 : 512) {  // 0, 512, 1024, 2048, ...
        if (memcmp(f.map + off, sig, 8) == 0) { found = true; break; }
    }
    if (!found) return fail("not an HDF5 file (signature not found)");
    const uint8_t* s = f.map + off;
    if (off + 96 > f.size) return fail("truncated superblock");
    const int version = s[8];
    if (version > 1) return fail("superblock version 2/3 (libver='latest') is not supported; CLAM/h5py default files are version 0");
    f.so = s[13];
    f.sl = s[14];
    if ((f.so != 8 && f.so != 4) || (f.sl != 8 && f.sl != 4)) return fail("unsupported size of offsets / lengths");
    const uint8_t* p = s + 24 + (version == 1 ? 4 : 0);
    const uint64_t base = Reader::le(p, f.so);
    p += 4 * f.so;  // base, free-space, end-of-file, driver-info addresses
    f.base = off;   // addresses are relative to the base address, which equals the user-block size
    if (base != H5_UNDEF && base != off && base != 0) f.base = base;
    // root group symbol table entry
    p += f.so;  // link name offset
    f.root_ohdr = Reader::le(p, f.so);
    p += f.so;
    const uint32_t cache = (uint32_t)Reader::le_raw(p, 4);
    p += 8;
    if (cache == 1) {
        f.root_btree = Reader::le(p, f.so);
        f.root_heap = Reader::le(p + f.so, f.so);
    }
    return MOC_OK;
}

// Walks the messages of a version-1 object header (following continuation blocks) and calls cb(type, data, size).
template <typename F>
int walk_object_header(Reader& r, uint64_t addr, F cb) {
    const uint8_t* h = r.at(addr, 16);
    if (!h) return fail("object header outside the file");
    if (memcmp(h, "OHDR", 4) == 0) return fail("version-2 object headers (libver='latest') are not supported");
    if (h[0] != 1) return fail("unknown object header version");
    int n_msgs = (int)Reader::le_raw(h + 2, 2);
    uint64_t block = addr + 16, block_len = Reader::le_raw(h + 8, 4);
    std::vector<std::pair<uint64_t, uint64_t>> pending;
    for (int guard = 0; guard < 64; ++guard) {
        uint64_t pos = 0;
        while (n_msgs > 0 && pos + 8 <= block_len) {
            const uint8_t* m = r.at(block + pos, 8);
            if (!m) return fail("object header message outside the file");
            const int type = (int)Reader::le_raw(m, 2), size = (int)Reader::le_raw(m + 2, 2);
            const uint8_t* d = r.at(block + pos + 8, size);
            if (!d && size) return fail("object header message outside the file");
            if (type == 0x0010) {
                if (size < r.f.so + r.f.sl) return fail("truncated object header continuation message");
                pending.push_back({Reader::le(d, r.f.so), Reader::le_raw(d + r.f.so, r.f.sl)});
            } else {
                const int rc = cb(type, d, size);
                if (rc != MOC_OK) return rc;
            }
            pos += 8 + (uint64_t)size;
            --n_msgs;
        }
        if (pending.empty() || n_msgs <= 0) break;
        block = pending.back().first;
        block_len = pending.back().second;
        pending.pop_back();
    }
    return MOC_OK;
}

// Finds `name` in a symbol-table group; returns its object header address.
int group_lookup(Reader& r, uint64_t btree, uint64_t heap, const char* name, uint64_t* out) {
    const uint8_t* hp = r.at(heap, 8 + 2 * r.f.sl + r.f.so);
    if (!hp || memcmp(hp, "HEAP", 4) != 0) return fail("local heap not found");
    const uint64_t heap_size = Reader::le_raw(hp + 8, r.f.sl);
    const uint64_t heap_data = Reader::le(hp + 8 + 2 * r.f.sl, r.f.so);
    std::vector<uint64_t> stack{btree};
    for (int guard = 0; !stack.empty() && guard < 100000; ++guard) {
        const uint64_t node = stack.back();
        stack.pop_back();
        const uint8_t* n = r.at(node, 8);
        if (!n) return fail("group node outside the file");
        if (memcmp(n, "TREE", 4) == 0) {
            if (n[4] != 0) return fail("group B-tree has the wrong node type");
            const int used = (int)Reader::le_raw(n + 6, 2);
            const uint8_t* e = r.at(node + 8 + 2 * r.f.so, (uint64_t)(2 * used + 1) * 8);
            if (!e) return fail("group B-tree node outside the file");
            for (int i = 0; i < used; ++i) stack.push_back(Reader::le(e + r.f.sl + (uint64_t)i * (r.f.sl + r.f.so), r.f.so));
        } else if (memcmp(n, "SNOD", 4) == 0) {
            const int n_sym = (int)Reader::le_raw(n + 6, 2);
            const int esz = 2 * r.f.so + 24;
            const uint8_t* e = r.at(node + 8, (uint64_t)n_sym * esz);
            if (!e) return fail("symbol table node outside the file");
            for (int i = 0; i < n_sym; ++i) {
                const uint64_t name_off = Reader::le_raw(e + (uint64_t)i * esz, r.f.so);
                if (name_off >= heap_size) continue;
                const char* s = reinterpret_cast<const char*>(r.at(heap_data + name_off, 1));
                if (!s) return fail("link name outside the file");
                const size_t maxlen = (size_t)(heap_size - name_off);
                if (strnlen(s, maxlen) < maxlen && strcmp(s, name) == 0) {
                    *out = Reader::le(e + (uint64_t)i * esz + r.f.so, r.f.so);
                    return MOC_OK;
                }
            }
        } else {
            return fail("unknown node in a group B-tree");
        }
    }
    set_error("moc_h5: no dataset named '%s' in the root group", name);
    return MOC_E_ARG;
}

int find_dataset(const H5File& f, const char* name, H5Dataset* ds) {
    Reader r(f);
    uint64_t btree = f.root_btree, heap = f.root_heap;
    if (btree == H5_UNDEF) {  // symbol table message of the root object header
        const int rc = walk_object_header(r, f.root_ohdr, [&](int type, const uint8_t* d, int size) {
            if (type == 0x0011 && size >= 2 * f.so) {
                btree = Reader::le(d, f.so);
                heap = Reader::le(d + f.so, f.so);
            }
            return (int)MOC_OK;
        });
        if (rc != MOC_OK) return rc;
        if (btree == H5_UNDEF) return fail("the root group is not a symbol-table group (new-style groups are not supported)");
    }
    uint64_t ohdr = H5_UNDEF;
    int rc = group_lookup(r, btree, heap, name, &ohdr);
    if (rc != MOC_OK) return rc;
    rc = walk_object_header(r, ohdr, [&](int type, const uint8_t* d, int size) -> int {
        if (type == 0x0001) {  // dataspace
            if (size < 4) return fail("short dataspace message");
            const int ver = d[0];
            ds->rank = d[1];
            if (ds->rank > 4) return fail("datasets of rank > 4 are not supported");
            const uint8_t* p = d + (ver == 1 ? 8 : 4);
            if ((p - d) + ds->rank * f.sl > size) return fail("short dataspace message");
            for (int i = 0; i < ds->rank; ++i) ds->dims[i] = Reader::le_raw(p + (uint64_t)i * f.sl, f.sl);
        } else if (type == 0x0003) {  // datatype
            if (size < 8) return fail("short datatype message");
            ds->type_class = d[0] & 15;
            ds->big_endian = d[1] & 1;
            ds->elem_size = (int)Reader::le_raw(d + 4, 4);
        } else if (type == 0x000B) {  // filter pipeline
            if (size >= 2 && d[1] > 0) ds->filtered = true;
        } else if (type == 0x0008) {  // data layout
            if (size < 2) return fail("short layout message");
            const int ver = d[0];
            if (ver == 3) {
                ds->layout = d[1];
                if (ds->layout == 0) {
                    ds->data_size = Reader::le_raw(d + 2, 2);
                    ds->compact = d + 4;
                } else if (ds->layout == 1) {
                    ds->data_addr = Reader::le(d + 2, f.so);
                    ds->data_size = Reader::le_raw(d + 2 + f.so, f.sl);
                } else if (ds->layout == 2) {
                    ds->chunk_rank = d[2];
                    if (ds->chunk_rank < 1 || ds->chunk_rank > 5) return fail("bad chunk dimensionality");
                    ds->chunk_btree = Reader::le(d + 3, f.so);
                    for (int i = 0; i < ds->chunk_rank; ++i) ds->chunk_dims[i] = (uint32_t)Reader::le_raw(d + 3 + f.so + 4 * i, 4);
                } else {
                    return fail("unknown layout class");
                }
            } else if (ver == 1 || ver == 2) {
                const int dim = d[1];
                ds->layout = d[2];
                if (dim < 1 || dim > 5) return fail("bad layout dimensionality");
                const uint8_t* p = d + 8;
                if (ds->layout != 0) {
                    const uint64_t a = Reader::le(p, f.so);
                    p += f.so;
                    if (ds->layout == 1) ds->data_addr = a; else ds->chunk_btree = a;
                }
                // `dim` size fields follow; for chunked storage that count already includes the trailing element-size
                // entry (it is one more than the dataspace rank)
                uint64_t total = 1;
                for (int i = 0; i < dim; ++i) {
                    ds->chunk_dims[i] = (uint32_t)Reader::le_raw(p + 4 * i, 4);
                    total *= ds->chunk_dims[i];
                }
                p += 4 * dim;
                if (ds->layout == 2) {
                    ds->chunk_rank = dim;
                } else if (ds->layout == 0) {
                    ds->data_size = Reader::le_raw(p, 4);
                    ds->compact = p + 4;
                } else {
                    ds->data_size = total;  // element count (unused: the dataspace gives the extent)
                }
            } else {
                return fail("layout message version 4 (libver='latest') is not supported");
            }
        }
        return MOC_OK;
    });
    if (rc != MOC_OK) return rc;
    if (ds->layout < 0 || ds->type_class < 0 || ds->elem_size <= 0) return fail("object is not a simple dataset");
    if (ds->filtered) return fail("filtered (compressed) datasets are not supported; CLAM writes none");
    if (ds->type_class > 1) return fail("only integer and floating-point datasets are supported");
    return MOC_OK;
}

uint64_t n_elems(const H5Dataset& ds) {
    uint64_t n = 1;
    for (int i = 0; i < ds.rank; ++i) n *= ds.dims[i];
    return n;
}

// Copies every allocated chunk of a chunked dataset into the dense row-major destination.
int read_chunked(const H5File& f, const H5Dataset& ds, uint8_t* dst) {
    Reader r(f);
    const int rank = ds.rank, es = ds.elem_size;
    if (ds.chunk_rank != rank + 1) return fail("chunk dimensionality does not match the dataspace");
    uint64_t chunk_bytes = es;
    for (int i = 0; i < rank; ++i) chunk_bytes *= ds.chunk_dims[i];
    const uint64_t key_size = 8 + 8ull * (rank + 1);
    // strides of the destination in elements
    uint64_t dstride[4] = {1, 1, 1, 1};
    for (int i = rank - 2; i >= 0; --i) dstride[i] = dstride[i + 1] * ds.dims[i + 1];
    memset(dst, 0, n_elems(ds) * es);
    if (ds.chunk_btree == H5_UNDEF) return MOC_OK;  // nothing allocated yet
    std::vector<uint64_t> stack{ds.chunk_btree};
    std::unordered_set<uint64_t> visited;   // a node reached twice means a cyclic (corrupt) tree: fail, do not spin
    while (!stack.empty()) {
        const uint64_t node = stack.back();
        stack.pop_back();
        if (!visited.insert(node).second) return fail("chunk B-tree is cyclic (corrupt file)");
        const uint8_t* n = r.at(node, 8 + 2 * f.so);
        if (!n || memcmp(n, "TREE", 4) != 0 || n[4] != 1) return fail("chunk B-tree node not found");
        const int level = n[5], used = (int)Reader::le_raw(n + 6, 2);
        const uint8_t* e = r.at(node + 8 + 2 * f.so, (uint64_t)used * (key_size + f.so) + key_size);
        if (!e) return fail("chunk B-tree node outside the file");
        for (int i = 0; i < used; ++i) {
            const uint8_t* key = e + (uint64_t)i * (key_size + f.so);
            const uint64_t child = Reader::le(key + key_size, f.so);
            if (level > 0) { stack.push_back(child); continue; }
            const uint32_t stored = (uint32_t)Reader::le_raw(key, 4), mask = (uint32_t)Reader::le_raw(key + 4, 4);
            if (mask != 0 || stored != chunk_bytes) return fail("filtered or odd-sized chunk");
            uint64_t off[4] = {0, 0, 0, 0};
            bool inside = true;
            for (int d = 0; d < rank; ++d) {
                off[d] = Reader::le_raw(key + 8 + 8ull * d, 8);
                if (off[d] >= ds.dims[d]) inside = false;
            }
            if (!inside) continue;  // chunk left over from a shrunk dataset
            const uint8_t* src = r.at(child, chunk_bytes);
            if (!src) return fail("chunk outside the file");
            // copy the intersection of the chunk with the dataset, innermost dimension contiguous
            uint64_t ext[4] = {1, 1, 1, 1}, cstride[4] = {1, 1, 1, 1};
            for (int d = 0; d < rank; ++d) {
                const uint64_t left = ds.dims[d] - off[d];
                ext[d] = left < ds.chunk_dims[d] ? left : ds.chunk_dims[d];
            }
            for (int d = rank - 2; d >= 0; --d) cstride[d] = cstride[d + 1] * ds.chunk_dims[d + 1];
            const int inner = rank - 1;
            uint64_t idx[4] = {0, 0, 0, 0};
            const uint64_t outer = (rank >= 1 ? 1 : 0) * (rank > 1 ? ext[0] : 1) * (rank > 2 ? ext[1] : 1) * (rank > 3 ? ext[2] : 1);
            for (uint64_t it = 0; it < outer; ++it) {
                uint64_t doff = 0, coff = 0;
                for (int d = 0; d < inner; ++d) {
                    doff += (off[d] + idx[d]) * dstride[d];
                    coff += idx[d] * cstride[d];
                }
                doff += off[inner];
                memcpy(dst + doff * es, src + coff * es, ext[inner] * es);
                for (int d = inner - 1; d >= 0; --d) {
                    if (++idx[d] < ext[d]) break;
                    idx[d] = 0;
                }
            }
        }
    }
    return MOC_OK;
}

}  // namespace
}  // namespace moc

using namespace moc;

extern "C" int moc_h5_open(const char* path, void** handle) {
    MOC_CHECK_ARG(path && handle, "moc_h5_open: null pointer");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_error("moc_h5_open: cannot open '%s'", path);
        return MOC_E_ARG;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 64) {
        close(fd);
        set_error("moc_h5_open: '%s' is too small to be an HDF5 file", path);
        return MOC_E_ARG;
    }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m != MAP_FAILED) madvise(m, (size_t)st.st_size, MADV_WILLNEED);   // start read-ahead of the whole bag now
    if (m == MAP_FAILED) {
        set_error("moc_h5_open: mmap of '%s' failed", path);
        return MOC_E_ARG;
    }
    H5File* f = new H5File();
    f->map = static_cast<const uint8_t*>(m);
    f->size = (size_t)st.st_size;
    const int rc = parse_superblock(*f);
    if (rc != MOC_OK) {
        munmap(m, f->size);
        delete f;
        return rc;
    }
    *handle = f;
    return MOC_OK;
}

extern "C" void moc_h5_close(void* handle) {
    if (!handle) return;
    H5File* f = static_cast<H5File*>(handle);
    munmap(const_cast<uint8_t*>(f->map), f->size);
    delete f;
}

extern "C" int moc_h5_dataset_info(void* handle, const char* name, int* rank, int64_t* dims4, int* type_class,
                                   int* elem_size) {
    MOC_CHECK_ARG(handle && name && rank && dims4 && type_class && elem_size, "moc_h5_dataset_info: null pointer");
    H5Dataset ds;
    const int rc = find_dataset(*static_cast<H5File*>(handle), name, &ds);
    if (rc != MOC_OK) return rc;
    *rank = ds.rank;
    for (int i = 0; i < 4; ++i) dims4[i] = i < ds.rank ? (int64_t)ds.dims[i] : 1;
    *type_class = ds.type_class;
    *elem_size = ds.elem_size;
    return MOC_OK;
}

extern "C" int moc_h5_read(void* handle, const char* name, void* dst_h, size_t dst_bytes) {
    MOC_CHECK_ARG(handle && name && dst_h, "moc_h5_read: null pointer");
    const H5File& f = *static_cast<H5File*>(handle);
    H5Dataset ds;
    int rc = find_dataset(f, name, &ds);
    if (rc != MOC_OK) return rc;
    if (ds.big_endian) return fail("big-endian datasets are not supported");
    const uint64_t bytes = n_elems(ds) * (uint64_t)ds.elem_size;
    if (dst_bytes < bytes) {
        set_error("moc_h5_read: destination holds %zu bytes, dataset '%s' needs %llu", dst_bytes, name, (unsigned long long)bytes);
        return MOC_E_WORKSPACE;
    }
    uint8_t* dst = static_cast<uint8_t*>(dst_h);
    Reader r(f);
    if (ds.layout == 0) {
        if (ds.data_size < bytes) return fail("compact dataset shorter than its dataspace");
        memcpy(dst, ds.compact, bytes);
        return MOC_OK;
    }
    if (ds.layout == 1) {
        if (ds.data_addr == H5_UNDEF) { memset(dst, 0, bytes); return MOC_OK; }
        const uint8_t* src = r.at(ds.data_addr, bytes);
        if (!src) return fail("contiguous dataset outside the file");
        memcpy(dst, src, bytes);
        return MOC_OK;
    }
    return read_chunked(f, ds, dst);
}
