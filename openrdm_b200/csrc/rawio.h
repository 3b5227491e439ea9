// rawio.h -- host-side reader of the MC-PDFT input container (BASELINE.json north_star
// "Data loading (libio)").
//
// The only grid + orbitals + RDM container the reference's ecosystem produces is the PySCF exporter's `data.h5`
// (/root/reference/src/libpyscf/libpyscf.py:613-690): groups /N (dimensions), /C (MO coefficients), /E (scalars),
// /D (dense RDMs), /GRIDS/{W,X,Y,Z}, /PHI/PHI_{AO,MO}{,_X,_Y,_Z} and the upper-packed COO RDMs under /SP_SYMM_D1,
// /SP_SYMM_D2 (libpyscf.py:246-463); libio's own files (opdm.h5 /D1a/D1a_DataSet ..., src/libio/HDF5_Read_Compact.cc:8-12)
// are a write-then-read round trip whose result mcpdft_energy discards (energy.cc:132-140).
//
// Two on-disk forms of that one schema are read here, both behind ordm_file_open():
//   * HDF5 itself, through the HDF5 C API, when the library is present at build time (__has_include(<hdf5.h>): it
//     is not in this image, so that branch is compiled out here);
//   * a raw little-endian container holding the same datasets under the same names -- the stand-in SURVEY.md names:
//       char     magic[8] = "ORDMRAW1"
//       uint64   n_datasets
//       uint64   toc_bytes                      (bytes of the table of contents that follows)
//       n_datasets x { uint32 name_len; char name[name_len]; uint32 dtype (0 f64, 1 i64, 2 i32); uint32 ndim;
//                      uint64 dims[ndim]; uint64 offset (from the start of the file, 64-byte aligned); uint64 nbytes }
//       data, row-major (C order) exactly as numpy / h5py hold the arrays
//     written by openrdm_b200/rawio.py (write_container; convert_h5 turns a data.h5 into it where h5py exists).
// The raw file is mmap'ed: ordm_file_dataset hands out pointers into the mapping, nothing is copied.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#if defined(__has_include)
#if __has_include(<hdf5.h>) && !defined(ORDM_NO_HDF5)
#define ORDM_WITH_HDF5 1
#include <hdf5.h>
#endif
#endif

namespace ordm {

enum RawDtype { RAW_F64 = 0, RAW_I64 = 1, RAW_I32 = 2 };

struct RawDataset {
  int dtype = RAW_F64, ndim = 0;
  uint64_t dims[4] = {0, 0, 0, 0};
  const void* data = nullptr;
  uint64_t nbytes = 0;
  uint64_t count() const { uint64_t n = 1; for (int i = 0; i < ndim; ++i) n *= dims[i]; return n; }
};

struct RawFile {
  std::map<std::string, RawDataset> sets;
  void* map = nullptr;
  size_t map_bytes = 0;
  int fd = -1;
  bool is_hdf5 = false;
#ifdef ORDM_WITH_HDF5
  hid_t h5 = -1;
  std::map<std::string, std::vector<char>> owned;   // datasets read from HDF5 live here
#endif
  std::string error;

  ~RawFile() { close_all(); }
  void close_all() {
    if (map && map != MAP_FAILED) munmap(map, map_bytes);
    map = nullptr;
    if (fd >= 0) ::close(fd);
    fd = -1;
#ifdef ORDM_WITH_HDF5
    if (h5 >= 0) H5Fclose(h5);
    h5 = -1;
#endif
  }

  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { error = std::string("cannot open ") + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { error = std::string(path) + ": too short for a container"; return false; }
    char magic[8];
    if (pread(fd, magic, 8, 0) != 8) { error = std::string(path) + ": unreadable"; return false; }
    if (std::memcmp(magic, "ORDMRAW1", 8) == 0) return open_raw(path, (size_t)st.st_size);
    if (std::memcmp(magic, "\x89HDF\r\n\x1a\n", 8) == 0) {
#ifdef ORDM_WITH_HDF5
      ::close(fd); fd = -1;
      h5 = H5Fopen(path, H5F_ACC_RDONLY, H5P_DEFAULT);
      if (h5 < 0) { error = std::string(path) + ": H5Fopen failed"; return false; }
      is_hdf5 = true;
      return true;
#else
      error = std::string(path) + " is an HDF5 file, but this library was built without HDF5 (<hdf5.h> not found at build time); "
              "convert it with openrdm_b200/rawio.py convert_h5 where h5py is available";
      return false;
#endif
    }
    error = std::string(path) + ": neither an ORDMRAW1 container nor an HDF5 file";
    return false;
  }

  bool open_raw(const char* path, size_t bytes) {
    map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { map = nullptr; error = std::string(path) + ": mmap failed"; return false; }
    map_bytes = bytes;
    const uint8_t* base = (const uint8_t*)map;
    uint64_t n = 0, toc = 0;
    std::memcpy(&n, base + 8, 8);
    std::memcpy(&toc, base + 16, 8);
    if (toc > bytes - 24 || n > (1u << 20)) { error = std::string(path) + ": corrupt table of contents"; return false; }
    const uint8_t* p = base + 24;
    const uint8_t* end = p + toc;
    auto take = [&](void* dst, size_t k) -> bool { if ((size_t)(end - p) < k) return false; std::memcpy(dst, p, k); p += k; return true; };
    for (uint64_t i = 0; i < n; ++i) {
      uint32_t name_len = 0, dtype = 0, ndim = 0;
      if (!take(&name_len, 4) || name_len > 4096 || (size_t)(end - p) < name_len) { error = std::string(path) + ": corrupt entry"; return false; }
      std::string name((const char*)p, name_len);
      p += name_len;
      RawDataset d;
      if (!take(&dtype, 4) || !take(&ndim, 4) || ndim > 4 || dtype > 2) { error = std::string(path) + ": corrupt entry " + name; return false; }
      d.dtype = (int)dtype; d.ndim = (int)ndim;
      for (uint32_t k = 0; k < ndim; ++k) if (!take(&d.dims[k], 8)) { error = std::string(path) + ": corrupt entry " + name; return false; }
      uint64_t off = 0;
      if (!take(&off, 8) || !take(&d.nbytes, 8)) { error = std::string(path) + ": corrupt entry " + name; return false; }
      const uint64_t elem = d.dtype == RAW_I32 ? 4 : 8;
      if (off > bytes || d.nbytes > bytes - off || d.nbytes != d.count() * elem) { error = std::string(path) + ": dataset " + name + " lies outside the file"; return false; }
      d.data = base + off;
      sets[name] = d;
    }
    return true;
  }

  // dataset by its data.h5 path; nullptr if absent
  const RawDataset* find(const std::string& name) {
    auto it = sets.find(name);
    if (it != sets.end()) return &it->second;
#ifdef ORDM_WITH_HDF5
    if (is_hdf5) return read_h5(name);
#endif
    return nullptr;
  }

#ifdef ORDM_WITH_HDF5
  const RawDataset* read_h5(const std::string& name) {
    if (H5Lexists(h5, name.c_str(), H5P_DEFAULT) <= 0) return nullptr;   // (intermediate groups of the schema's paths always exist when the leaf does)
    hid_t ds = H5Dopen2(h5, name.c_str(), H5P_DEFAULT);
    if (ds < 0) return nullptr;
    RawDataset d;
    hid_t sp = H5Dget_space(ds), ty = H5Dget_type(ds);
    const int nd = H5Sget_simple_extent_ndims(sp);
    hsize_t dims[8] = {0};
    bool ok = nd >= 0 && nd <= 4;
    if (ok && nd > 0) H5Sget_simple_extent_dims(sp, dims, nullptr);
    d.ndim = nd;
    for (int i = 0; i < nd && ok; ++i) d.dims[i] = (uint64_t)dims[i];
    const bool is_int = H5Tget_class(ty) == H5T_INTEGER;
    d.dtype = is_int ? RAW_I64 : RAW_F64;
    d.nbytes = d.count() * 8;
    std::vector<char>& buf = owned[name];
    buf.resize((size_t)std::max<uint64_t>(d.nbytes, 8));
    if (ok) ok = H5Dread(ds, is_int ? H5T_NATIVE_INT64 : H5T_NATIVE_DOUBLE, H5S_ALL, H5S_ALL, H5P_DEFAULT, buf.data()) >= 0;
    H5Tclose(ty); H5Sclose(sp); H5Dclose(ds);
    if (!ok) { owned.erase(name); return nullptr; }
    d.data = buf.data();
    sets[name] = d;
    return &sets[name];
  }
#endif

  // a scalar (0-d or one-element) dataset as double / integer
  bool scalar(const std::string& name, double& out) {
    const RawDataset* d = find(name);
    if (!d || d->count() != 1) return false;
    if (d->dtype == RAW_F64) out = *(const double*)d->data;
    else if (d->dtype == RAW_I64) out = (double)*(const int64_t*)d->data;
    else out = (double)*(const int32_t*)d->data;
    return true;
  }
  // an index / count array as int (COO index lists are int64 from numpy, int32 from elsewhere)
  bool ints(const std::string& name, std::vector<int>& out) {
    const RawDataset* d = find(name);
    if (!d || d->dtype == RAW_F64) return false;
    const uint64_t n = d->count();
    out.resize(n);
    for (uint64_t i = 0; i < n; ++i) out[i] = d->dtype == RAW_I64 ? (int)((const int64_t*)d->data)[i] : ((const int32_t*)d->data)[i];
    return true;
  }
};

}  // namespace ordm
