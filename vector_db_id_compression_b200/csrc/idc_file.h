// idc_file.h -- the flat on-disk form of a blob: a 32-byte header and a run of sections, each an
// 8-byte length followed by the bytes padded to a multiple of 8. What is stored is exactly what the export entry
// points return (the wire form a gather over NCCL carries), so a file written on one box loads on any other; derived
// arrays (select samples, chunk directories, decode plans) are rebuilt by the import entry points.
//
//   offset 0   char[8]  "IDCBLOB\0"
//          8   u32      format version (1)
//         12   u32      kind: 1 ROC, 2 Elias-Fano, 3 wavelet tree
//         16   u64      number of sections
//         24   u64      reserved (0)
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "idc_host.h"

namespace idc {

constexpr uint32_t kFileVersion = 1;
enum FileKind : uint32_t { kFileRoc = 1, kFileEf = 2, kFileWt = 3 };

struct FileWriter {
    FILE* f = nullptr;
    bool ok = true;
    ~FileWriter() {
        if (f) fclose(f);
    }
    int open(const char* path, uint32_t kind, uint64_t nsections) {
        f = fopen(path, "wb");
        IDC_REQUIRE(f != nullptr, IDC_ERR_ARG, "cannot open %s for writing", path);
        char magic[8] = {'I', 'D', 'C', 'B', 'L', 'O', 'B', 0};
        const uint64_t reserved = 0;
        ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&kFileVersion, 4, 1, f) == 1 && fwrite(&kind, 4, 1, f) == 1 &&
             fwrite(&nsections, 8, 1, f) == 1 && fwrite(&reserved, 8, 1, f) == 1;
        return IDC_OK;
    }
    void section(const void* p, uint64_t bytes) {
        static const char zero[8] = {0};
        ok = ok && fwrite(&bytes, 8, 1, f) == 1;
        if (bytes) ok = ok && fwrite(p, 1, bytes, f) == bytes;
        if (bytes & 7) ok = ok && fwrite(zero, 1, 8 - (bytes & 7), f) == 8 - (bytes & 7);
    }
    template <typename T>
    void vec(const std::vector<T>& v) { section(v.data(), v.size() * sizeof(T)); }
    int close(const char* path) {
        ok = ok && fclose(f) == 0;
        f = nullptr;
        IDC_REQUIRE(ok, IDC_ERR_ARG, "write to %s failed", path);
        return IDC_OK;
    }
};

struct FileReader {
    FILE* f = nullptr;
    uint64_t nsections = 0, next = 0;
    ~FileReader() {
        if (f) fclose(f);
    }
    int open(const char* path, uint32_t kind) {
        f = fopen(path, "rb");
        IDC_REQUIRE(f != nullptr, IDC_ERR_ARG, "cannot open %s", path);
        char magic[8];
        uint32_t ver = 0, k = 0;
        uint64_t reserved = 0;
        const bool got = fread(magic, 1, 8, f) == 8 && fread(&ver, 4, 1, f) == 1 && fread(&k, 4, 1, f) == 1 &&
                         fread(&nsections, 8, 1, f) == 1 && fread(&reserved, 8, 1, f) == 1;
        IDC_REQUIRE(got && memcmp(magic, "IDCBLOB", 8) == 0, IDC_ERR_ARG, "%s is not a blob file", path);
        IDC_REQUIRE(ver == kFileVersion, IDC_ERR_ARG, "%s: format version %u, this library reads %u", path, ver, kFileVersion);
        IDC_REQUIRE(k == kind, IDC_ERR_ARG, "%s holds a blob of kind %u, expected %u", path, k, kind);
        return IDC_OK;
    }
    // next section into a vector of T (its length must be a whole number of elements)
    template <typename T>
    int vec(std::vector<T>& v) {
        uint64_t bytes = 0;
        IDC_REQUIRE(next < nsections && fread(&bytes, 8, 1, f) == 1, IDC_ERR_ARG, "blob file truncated (section %llu)",
                    (unsigned long long)next);
        IDC_REQUIRE(bytes % sizeof(T) == 0, IDC_ERR_ARG, "blob file: section %llu has %llu bytes, not a multiple of %zu",
                    (unsigned long long)next, (unsigned long long)bytes, sizeof(T));
        v.resize(bytes / sizeof(T));
        if (bytes) IDC_REQUIRE(fread(v.data(), 1, bytes, f) == bytes, IDC_ERR_ARG, "blob file truncated inside section %llu", (unsigned long long)next);
        if (bytes & 7) {
            char pad[8];
            IDC_REQUIRE(fread(pad, 1, 8 - (bytes & 7), f) == 8 - (bytes & 7), IDC_ERR_ARG, "blob file truncated (padding)");
        }
        next++;
        return IDC_OK;
    }
};

}  // namespace idc
