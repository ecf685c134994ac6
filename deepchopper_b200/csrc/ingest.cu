// FASTQ ingest, file level (no GPU work in this file): read a plain / gzip / BGZF file into one host buffer, inflating
// BGZF blocks on host threads.
//
// Reference code replaced: the compression sniffing and readers of src/output/writefq.rs:84-193 (plain, gzip and bgzip
// FASTQ through noodles / flate2, one thread) and the file access of deepchopper/data/only_fq.py:21-85 (pyfastx).  The
// record boundaries are found afterwards by the newline index (deepchopper_b200/encode.py index_fastq) and the
// bytes -> token / quality work runs on the GPU (dcb200_encode_batch).
//
// A BGZF file (what `deepchopper chop` writes and what bgzip produces) is a chain of independent <= 64 KiB gzip
// members whose compressed size sits in the "BC" extra field and whose inflated size in the trailer, so every block's
// place in the output is known before anything is inflated and the blocks inflate in parallel.  A plain gzip stream
// has no such index and is inflated by one thread.
#include "common.cuh"

#include <zlib.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace {

struct Block {
  size_t in_off;    // start of the raw deflate data
  uint32_t in_len;  // compressed bytes
  size_t out_off;
  uint32_t out_len;
  uint32_t crc;
};

// parses the member at `pos`; returns false if it is not a well-formed BGZF block
bool parse_bgzf(const uint8_t* d, size_t n, size_t pos, Block* b, size_t* next) {
  if (pos + 18 > n) return false;
  if (d[pos] != 0x1f || d[pos + 1] != 0x8b || d[pos + 2] != 8 || !(d[pos + 3] & 4)) return false;
  const uint32_t xlen = d[pos + 10] | (d[pos + 11] << 8);
  size_t x = pos + 12;
  const size_t xend = x + xlen;
  if (xend > n) return false;
  int bsize = -1;
  while (x + 4 <= xend) {
    const uint32_t slen = d[x + 2] | (d[x + 3] << 8);
    if (d[x] == 'B' && d[x + 1] == 'C' && slen == 2 && x + 6 <= xend) bsize = (d[x + 4] | (d[x + 5] << 8)) + 1;
    x += 4 + slen;
  }
  if (bsize < 0 || pos + (size_t)bsize > n || (size_t)bsize < 12 + xlen + 8) return false;
  if (d[pos + 3] & ~4) return false;  // other header flags (name, comment, hcrc) are not BGZF
  b->in_off = xend;
  b->in_len = (uint32_t)(bsize - (12 + xlen) - 8);
  const uint8_t* t = d + pos + bsize - 8;
  b->crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
  b->out_len = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
  *next = pos + bsize;
  return true;
}

int inflate_stream(const uint8_t* d, size_t n, std::vector<uint8_t>& out) {  // plain gzip, possibly several members
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) return -1;
  out.resize(n * 4 + (1 << 16));
  size_t have = 0;
  zs.next_in = const_cast<Bytef*>(d);
  size_t in_left = n;
  for (;;) {
    if (have == out.size()) out.resize(out.size() * 2);
    const size_t room = out.size() - have;
    zs.next_out = out.data() + have;
    zs.avail_out = (uInt)std::min<size_t>(room, 1u << 30);
    zs.avail_in = (uInt)std::min<size_t>(in_left, 1u << 30);
    const uInt in_before = zs.avail_in, out_before = zs.avail_out;
    const int rc = inflate(&zs, Z_NO_FLUSH);
    in_left -= in_before - zs.avail_in;
    have += out_before - zs.avail_out;
    if (rc == Z_STREAM_END) {
      if (in_left == 0) break;
      if (inflateReset(&zs) != Z_OK) {  // next member
        inflateEnd(&zs);
        return -1;
      }
      continue;
    }
    if (rc != Z_OK && rc != Z_BUF_ERROR) {
      inflateEnd(&zs);
      return -1;
    }
    if (rc == Z_BUF_ERROR && in_left == 0 && zs.avail_out != 0) {  // truncated stream
      inflateEnd(&zs);
      return -1;
    }
  }
  inflateEnd(&zs);
  out.resize(have);
  return 0;
}

}  // namespace

extern "C" void dcb200_free(void* p) { free(p); }

extern "C" int dcb200_read_file_inflate(const char* path, int32_t threads, uint8_t** out, int64_t* out_len,
                                        int32_t* kind) {
  using dcb::set_error;
  if (!path || !out || !out_len) {
    set_error("dcb200_read_file_inflate: null argument");
    return DCB200_EINVAL;
  }
  *out = nullptr;
  *out_len = 0;
  if (kind) *kind = 0;
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error("dcb200_read_file_inflate: cannot open '%s'", path);
    return DCB200_EINVAL;
  }
  fseek(f, 0, SEEK_END);
  const long long fsz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> raw((size_t)std::max(0LL, fsz));
  if (fsz > 0 && fread(raw.data(), 1, raw.size(), f) != raw.size()) {
    fclose(f);
    set_error("dcb200_read_file_inflate: short read of '%s'", path);
    return DCB200_EINVAL;
  }
  fclose(f);
  auto give = [&](const uint8_t* src, size_t n) -> int {
    uint8_t* p = static_cast<uint8_t*>(malloc(n ? n : 1));
    if (!p) {
      set_error("dcb200_read_file_inflate: out of memory (%zu bytes)", n);
      return DCB200_ENOMEM;
    }
    if (n) memcpy(p, src, n);
    *out = p;
    *out_len = (int64_t)n;
    return DCB200_OK;
  };
  if (raw.size() < 2 || raw[0] != 0x1f || raw[1] != 0x8b) return give(raw.data(), raw.size());  // plain text

  // BGZF? every member must parse as a block
  std::vector<Block> blocks;
  size_t pos = 0, total = 0;
  bool bgzf = true;
  while (pos < raw.size()) {
    Block b;
    size_t next;
    if (!parse_bgzf(raw.data(), raw.size(), pos, &b, &next)) {
      bgzf = false;
      break;
    }
    b.out_off = total;
    total += b.out_len;
    blocks.push_back(b);
    pos = next;
  }
  if (!bgzf) {
    std::vector<uint8_t> o;
    if (inflate_stream(raw.data(), raw.size(), o) != 0) {
      set_error("dcb200_read_file_inflate: '%s' is not a valid gzip stream", path);
      return DCB200_EINVAL;
    }
    if (kind) *kind = 1;
    return give(o.data(), o.size());
  }
  if (kind) *kind = 2;
  uint8_t* dst = static_cast<uint8_t*>(malloc(total ? total : 1));
  if (!dst) {
    set_error("dcb200_read_file_inflate: out of memory (%zu bytes)", total);
    return DCB200_ENOMEM;
  }
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min<int>(T, 256));
  T = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, blocks.size()));
  std::atomic<size_t> next_block{0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) {
      bad = 1;
      return;
    }
    for (;;) {
      const size_t first = next_block.fetch_add(16);
      if (first >= blocks.size() || bad) break;
      for (size_t i = first; i < std::min(blocks.size(), first + 16); ++i) {
        const Block& b = blocks[i];
        inflateReset(&zs);
        zs.next_in = raw.data() + b.in_off;
        zs.avail_in = b.in_len;
        zs.next_out = dst + b.out_off;
        zs.avail_out = b.out_len;
        const int rc = inflate(&zs, Z_FINISH);
        if (!((rc == Z_STREAM_END || (rc == Z_BUF_ERROR && b.out_len == 0)) && zs.avail_out == 0) ||
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + b.out_off, b.out_len) != b.crc) {
          bad = 1;
          break;
        }
      }
    }
    inflateEnd(&zs);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  if (bad) {
    free(dst);
    set_error("dcb200_read_file_inflate: corrupt BGZF block in '%s'", path);
    return DCB200_EINVAL;
  }
  *out = dst;
  *out_len = (int64_t)total;
  return DCB200_OK;
}
