// Minimal HDF5 writer (see hemo_h5.h).  Structures follow the public "HDF5 File Format Specification
// Version 1.1/2.0" (superblock v0, object header v1, B-tree v1, local heap, symbol table nodes).
#include "hemo_h5.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>

namespace hemo { namespace h5 {

namespace {
const uint64_t UNDEF = ~0ull;
const int GROUP_INTERNAL_K = 16;      // superblock field; the group B-tree has one node with one child
const int ISTORE_K = 32;              // chunk B-tree K: fixed default for superblock v0 files

struct Buf {
  std::vector<uint8_t> b;
  void u8(unsigned v) { b.push_back((uint8_t)v); }
  void u16(unsigned v) { u8(v & 255); u8((v >> 8) & 255); }
  void u32(uint64_t v) { for (int i = 0; i < 4; i++) u8((unsigned)((v >> (8*i)) & 255)); }
  void u64(uint64_t v) { for (int i = 0; i < 8; i++) u8((unsigned)((v >> (8*i)) & 255)); }
  void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void pad8() { while (b.size() % 8) b.push_back(0); }
  size_t size() const { return b.size(); }
};

// datatype message body (v1), not padded
void put_datatype(Buf& o, Type t) {
  switch (t) {
    case F32: case F64: {
      const bool d = (t == F64);
      o.u8(0x11); o.u8(0x20); o.u8(d ? 63 : 31); o.u8(0); o.u32(d ? 8 : 4);       // class 1 (float), LE, implied-msb mantissa, sign bit
      o.u16(0); o.u16(d ? 64 : 32); o.u8(d ? 52 : 23); o.u8(d ? 11 : 8); o.u8(0); o.u8(d ? 52 : 23); o.u32(d ? 1023 : 127);
      break; }
    case I32: case I64: {
      const bool l = (t == I64);
      o.u8(0x10); o.u8(0x08); o.u8(0); o.u8(0); o.u32(l ? 8 : 4);                  // class 0 (fixed point), LE, signed
      o.u16(0); o.u16(l ? 64 : 32);
      break; }
  }
}
// simple dataspace message body (v1, no max dims)
void put_dataspace(Buf& o, const std::vector<uint64_t>& dims) {
  o.u8(1); o.u8((unsigned)dims.size()); o.u8(0); o.u8(0); o.u32(0);
  for (uint64_t d : dims) o.u64(d);
}
// one v1 header message: type, body (padded to 8)
void put_message(Buf& o, unsigned type, const Buf& body, unsigned flags = 0) {
  const size_t padded = (body.size() + 7) / 8 * 8;
  o.u16(type); o.u16((unsigned)padded); o.u8(flags); o.u8(0); o.u8(0); o.u8(0);
  o.bytes(body.b.data(), body.size()); o.zeros(padded - body.size());
}
// v1 object header around `n` messages already serialised in msgs
Buf object_header(const Buf& msgs, unsigned n) {
  Buf h;
  h.u8(1); h.u8(0); h.u16(n); h.u32(1); h.u32(msgs.size()); h.u32(0);                // prefix is 16 bytes (12 + alignment pad)
  h.bytes(msgs.b.data(), msgs.size());
  return h;
}
}  // namespace

size_t type_size(Type t) { return (t == F64 || t == I64) ? 8 : 4; }

Writer::Writer(const std::string& p, int deflate_level) : path(p), level(deflate_level) {
  fp = fopen(p.c_str(), "wb");
  if (!fp) { fail("cannot open " + p); return; }
  uint8_t z[96]; memset(z, 0, sizeof(z));
  if (fwrite(z, 1, 96, fp) != 96) fail("write failed");                             // the superblock goes here at close()
  eof = 96;
}
Writer::~Writer() { if (fp) close(); }
void Writer::fail(const std::string& m) { if (!failed) err = m; failed = true; }

uint64_t Writer::append(const void* p, size_t n) {
  if (!fp || failed) return UNDEF;
  static const uint8_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (eof % 8) { const size_t k = 8 - eof % 8; if (fwrite(z, 1, k, fp) != k) fail("write failed"); eof += k; }
  const uint64_t at = eof;
  if (n && fwrite(p, 1, n, fp) != n) fail("write failed");
  eof += n;
  return at;
}

void Writer::attribute(const std::string& name, Type t, const void* data, size_t n) {
  Attr a; a.name = name; a.t = t; a.n = n;
  a.data.assign((const uint8_t*)data, (const uint8_t*)data + n*type_size(t));
  for (auto& e : attrs) if (e.name == name) { e = a; return; }                       // H5LTset_attribute_* overwrites
  attrs.push_back(a);
}

void Writer::dataset(const std::string& name, Type t, const std::vector<uint64_t>& dims, const void* data,
                     const std::vector<uint64_t>& chunk_in) {
  if (!ok()) return;
  for (auto& e : dsets) if (e.name == name) { fail("duplicate dataset " + name); return; }
  Dset d; d.name = name; d.t = t; d.dims = dims;
  const size_t es = type_size(t), r = dims.size();
  uint64_t total = 1; for (uint64_t v : dims) total *= v;
  d.chunked = (level >= 0) && chunk_in.size() == r && r > 0;
  if (!d.chunked) {
    d.data_size = total*es;
    d.data_addr = total ? append(data, (size_t)d.data_size) : UNDEF;
    dsets.push_back(d);
    return;
  }
  d.deflated = true; d.chunk = chunk_in;
  for (auto& c : d.chunk) if (c < 1) c = 1;
  // walk the chunk grid in row-major order (= ascending B-tree key order)
  std::vector<uint64_t> nch(r), idx(r, 0), stride(r);
  uint64_t nchunks = 1, celems = 1;
  for (size_t k = 0; k < r; k++) { nch[k] = (dims[k] + d.chunk[k] - 1)/d.chunk[k]; nchunks *= nch[k]; celems *= d.chunk[k]; }
  for (size_t k = r; k-- > 0;) stride[k] = (k + 1 == r) ? 1 : stride[k + 1]*dims[k + 1];
  if (celems*es > 0xFFFFFFFFull) { fail("chunk larger than 4 GiB in " + name); return; }
  std::vector<uint8_t> raw((size_t)(celems*es)), comp(compressBound((uLong)(celems*es)));
  const uint8_t* src = (const uint8_t*)data;
  for (uint64_t c = 0; c < nchunks && total; c++) {
    std::fill(raw.begin(), raw.end(), 0);
    // copy the rows (runs along the last dimension) of this chunk that lie inside the dataset
    const uint64_t last0 = idx[r - 1]*d.chunk[r - 1];
    const uint64_t run = std::min<uint64_t>(d.chunk[r - 1], dims[r - 1] - last0);
    std::vector<uint64_t> in(r, 0);
    const uint64_t rows = celems/d.chunk[r - 1];
    for (uint64_t row = 0; row < rows; row++) {
      uint64_t rem = row, off = last0; bool inside = true;
      for (size_t k = r - 1; k-- > 0;) {
        in[k] = rem % d.chunk[k]; rem /= d.chunk[k];
        const uint64_t g = idx[k]*d.chunk[k] + in[k];
        if (g >= dims[k]) { inside = false; break; }
        off += g*stride[k];
      }
      if (inside) memcpy(&raw[(size_t)(row*d.chunk[r - 1]*es)], src + off*es, (size_t)(run*es));
    }
    uLongf clen = (uLongf)comp.size();
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), level) != Z_OK) { fail("deflate failed"); return; }
    ChunkRec cr; cr.nbytes = (uint32_t)clen; cr.addr = append(comp.data(), clen);
    cr.offset.resize(r); for (size_t k = 0; k < r; k++) cr.offset[k] = idx[k]*d.chunk[k];
    d.chunks.push_back(cr);
    for (size_t k = r; k-- > 0;) { if (++idx[k] < nch[k]) break; idx[k] = 0; }
  }
  d.btree_addr = write_chunk_btree(d);
  dsets.push_back(d);
}

// B-tree v1, node type 1 (raw data chunks).  Key = {chunk bytes u32, filter mask u32, offsets u64 x (rank+1)}.
uint64_t Writer::write_chunk_btree(Dset& d) {
  if (d.chunks.empty()) return UNDEF;
  const size_t r = d.dims.size(), keysize = 8 + 8*(r + 1), cap = 2*ISTORE_K;
  const size_t nodesize = 24 + (cap + 1)*keysize + cap*8;
  struct Ent { uint32_t nbytes; std::vector<uint64_t> off; uint64_t child; };
  std::vector<Ent> level_e;
  for (auto& c : d.chunks) level_e.push_back({c.nbytes, c.offset, c.addr});
  // the key that closes the right-most node of every level: one chunk row past the end
  Ent last; last.nbytes = 0; last.child = 0; last.off.assign(r, 0);
  last.off[0] = d.chunks.back().offset[0] + d.chunk[0];
  for (int lvl = 0;; lvl++) {
    const size_t nn = (level_e.size() + cap - 1)/cap;
    if (eof % 8) append(nullptr, 0);
    uint64_t base = (eof + 7)/8*8;
    std::vector<Ent> up;
    for (size_t n = 0; n < nn; n++) {
      const size_t a = n*cap, b = std::min(level_e.size(), a + cap);
      Buf o;
      o.bytes("TREE", 4); o.u8(1); o.u8((unsigned)lvl); o.u16((unsigned)(b - a));
      o.u64(n ? base + (n - 1)*nodesize : UNDEF); o.u64(n + 1 < nn ? base + (n + 1)*nodesize : UNDEF);
      for (size_t e = a; e <= b; e++) {
        const Ent& k = (e < level_e.size()) ? level_e[e] : last;
        o.u32(k.nbytes); o.u32(0);
        for (size_t q = 0; q < r; q++) o.u64(k.off[q]);
        o.u64(0);
        if (e < b) o.u64(level_e[e].child);
      }
      o.zeros(nodesize - o.size());
      const uint64_t at = append(o.b.data(), o.size());
      up.push_back({level_e[a].nbytes, level_e[a].off, at});
    }
    if (nn == 1) return up[0].child;
    level_e.swap(up);
  }
}

bool Writer::close() {
  if (!fp) return false;
  if (!failed) {
    // ---- dataset object headers
    std::sort(dsets.begin(), dsets.end(), [](const Dset& a, const Dset& b) { return a.name < b.name; });
    for (auto& d : dsets) {
      Buf msgs; unsigned n = 0;
      { Buf b; put_dataspace(b, d.dims); put_message(msgs, 0x0001, b, 1); n++; }
      { Buf b; put_datatype(b, d.t); put_message(msgs, 0x0003, b, 1); n++; }
      { Buf b; b.u8(2); b.u8(d.chunked ? 3 : 2); b.u8(2); b.u8(1); b.u32(0); put_message(msgs, 0x0005, b, 1); n++; }   // fill value v2: default fill
      if (d.deflated) {
        Buf b; b.u8(1); b.u8(1); b.zeros(6);
        b.u16(1); b.u16(0); b.u16(1); b.u16(1); b.u32((uint64_t)level); b.u32(0);     // deflate, optional, 1 client value + pad
        put_message(msgs, 0x000B, b, 1); n++;
      }
      { Buf b; b.u8(3);
        if (d.chunked) {
          b.u8(2); b.u8((unsigned)d.dims.size() + 1); b.u64(d.btree_addr);
          for (uint64_t c : d.chunk) b.u32(c);
          b.u32(type_size(d.t));
        } else { b.u8(1); b.u64(d.data_addr); b.u64(d.data_size); }
        put_message(msgs, 0x0008, b); n++; }
      Buf h = object_header(msgs, n);
      d.header_addr = append(h.b.data(), h.size());
    }
    // ---- root group: local heap with the link names, one symbol table node, one B-tree node
    Buf heap; heap.zeros(8);                                                         // offset 0 = ""
    std::vector<uint64_t> name_off;
    for (auto& d : dsets) { name_off.push_back(heap.size()); heap.bytes(d.name.c_str(), d.name.size() + 1); heap.pad8(); }
    const uint64_t heap_data_addr = append(heap.b.data(), heap.size());
    Buf hh; hh.bytes("HEAP", 4); hh.u8(0); hh.zeros(3); hh.u64(heap.size()); hh.u64(1 /* H5HL_FREE_NULL: no free block */); hh.u64(heap_data_addr);
    const uint64_t heap_addr = append(hh.b.data(), hh.size());
    unsigned leaf_k = 4; while (2*leaf_k < dsets.size()) leaf_k *= 2;
    Buf sn; sn.bytes("SNOD", 4); sn.u8(1); sn.u8(0); sn.u16((unsigned)dsets.size());
    for (size_t i = 0; i < dsets.size(); i++) { sn.u64(name_off[i]); sn.u64(dsets[i].header_addr); sn.u32(0); sn.u32(0); sn.zeros(16); }
    sn.zeros(8 + 2*leaf_k*40 - sn.size());
    const uint64_t snod_addr = append(sn.b.data(), sn.size());
    Buf bt; bt.bytes("TREE", 4); bt.u8(0); bt.u8(0); bt.u16(dsets.empty() ? 0 : 1); bt.u64(UNDEF); bt.u64(UNDEF);
    bt.u64(0);
    if (!dsets.empty()) { bt.u64(snod_addr); bt.u64(name_off.back()); }
    bt.zeros(24 + (2*GROUP_INTERNAL_K + 1)*8 + 2*GROUP_INTERNAL_K*8 - bt.size());
    const uint64_t btree_addr = append(bt.b.data(), bt.size());
    // ---- root object header: symbol table message + attributes
    Buf msgs; unsigned n = 0;
    { Buf b; b.u64(btree_addr); b.u64(heap_addr); put_message(msgs, 0x0011, b); n++; }
    for (auto& a : attrs) {
      Buf dt, ds; put_datatype(dt, a.t); put_dataspace(ds, std::vector<uint64_t>{(uint64_t)a.n});
      Buf b; b.u8(1); b.u8(0); b.u16((unsigned)a.name.size() + 1); b.u16((unsigned)dt.size()); b.u16((unsigned)ds.size());
      b.bytes(a.name.c_str(), a.name.size() + 1); b.pad8();
      b.bytes(dt.b.data(), dt.size()); b.pad8();
      b.bytes(ds.b.data(), ds.size()); b.pad8();
      b.bytes(a.data.data(), a.data.size());
      put_message(msgs, 0x000C, b); n++;
    }
    Buf rh = object_header(msgs, n);
    const uint64_t root_addr = append(rh.b.data(), rh.size());
    append(nullptr, 0);
    // ---- superblock v0
    Buf sb; static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    sb.bytes(sig, 8); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(8); sb.u8(8); sb.u8(0);
    sb.u16(leaf_k); sb.u16(GROUP_INTERNAL_K); sb.u32(0);
    sb.u64(0); sb.u64(UNDEF); sb.u64(eof); sb.u64(UNDEF);
    sb.u64(0); sb.u64(root_addr); sb.u32(1); sb.u32(0); sb.u64(btree_addr); sb.u64(heap_addr);
    if (fseek(fp, 0, SEEK_SET) != 0 || fwrite(sb.b.data(), 1, sb.size(), fp) != sb.size()) fail("superblock write failed");
  }
  if (fclose(fp) != 0) fail("close failed");
  fp = nullptr;
  return !failed;
}

}}  // namespace hemo::h5
