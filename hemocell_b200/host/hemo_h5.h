// hemo_h5 -- a minimal stand-alone HDF5 (file format spec 1.1 / "v0 superblock") writer.
//
// The reference writes its field and particle output through libhdf5 + H5LT
// (io/ParticleHdf5IO.cpp:60-194, io/FluidHdf5IO.hh:36-211): one file per atomic block, a flat root
// group of N-d datasets (chunked, deflate 7) and a handful of numeric root attributes.  libhdf5 is
// not part of this image, so this writer produces exactly that subset of the format by hand:
//   superblock v0, root group = v1 object header + symbol-table message (one B-tree v1 node, one
//   SNOD, one local heap), per dataset a v1 object header (dataspace v1, datatype v1, fill value v2,
//   layout v3 [contiguous | chunked + B-tree v1 type 1], filter pipeline v1 [deflate]), attribute
//   messages v1 on the root header.
// Raw data is streamed to the file as datasets are added; metadata is laid out at close().
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace hemo { namespace h5 {

enum Type { F32 = 0, F64 = 1, I32 = 2, I64 = 3 };
size_t type_size(Type t);

class Writer {
 public:
  // deflate_level < 0: contiguous, uncompressed datasets.  Otherwise chunked + deflate like the reference.
  explicit Writer(const std::string& path, int deflate_level = 7);
  ~Writer();
  bool ok() const { return fp != nullptr && !failed; }
  // H5LTset_attribute_{double,long,int,float}(file, "/", name, data, n): 1-d attribute of n values on "/"
  void attribute(const std::string& name, Type t, const void* data, size_t n);
  // H5Dcreate2 + H5Dwrite of a whole dataset.  chunk may be empty (-> contiguous); row-major data.
  void dataset(const std::string& name, Type t, const std::vector<uint64_t>& dims, const void* data,
               const std::vector<uint64_t>& chunk = std::vector<uint64_t>());
  bool close();
  const std::string& error() const { return err; }

 private:
  struct Attr { std::string name; Type t; std::vector<uint8_t> data; size_t n; };
  struct ChunkRec { uint64_t addr; uint32_t nbytes; std::vector<uint64_t> offset; };
  struct Dset {
    std::string name; Type t; std::vector<uint64_t> dims, chunk;
    uint64_t data_addr = 0, data_size = 0;       // contiguous
    std::vector<ChunkRec> chunks; bool chunked = false; bool deflated = false;
    uint64_t btree_addr = 0, header_addr = 0;
  };
  uint64_t append(const void* p, size_t n);      // at the 8-byte aligned end of file, returns the address
  uint64_t write_chunk_btree(Dset& d);
  void fail(const std::string& m);
  FILE* fp = nullptr; bool failed = false; std::string err, path;
  int level;
  uint64_t eof = 0;
  std::vector<Attr> attrs;
  std::vector<Dset> dsets;
};

}}  // namespace hemo::h5
