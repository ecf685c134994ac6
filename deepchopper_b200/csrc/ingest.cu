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

// ---- record index ------------------------------------------------------------------------------------------------------
// Record boundaries of a FASTQ text (4-line records), found on host threads: every thread counts the newlines of its
// slice of the buffer (memchr), a prefix sum places the slices' lines, every thread writes its line starts, then records
// are assembled and validated in parallel.  Semantics of the reference's reader (deepchopper/data/only_fq.py:21-85 via
// pyfastx, noodles in src/output/writefq.rs): '\r' before '\n' is stripped, trailing empty lines are ignored, the id is
// the header up to the first blank, sequence and quality must be non-empty and of equal length.
extern "C" int dcb200_index_fastq(const uint8_t* fastq, int64_t n_bytes, int32_t threads, dcb200_fastq_index_arrays* out) {
  using dcb::set_error;
  if ((!fastq && n_bytes) || n_bytes < 0 || !out) {
    set_error("dcb200_index_fastq: bad argument");
    return DCB200_EINVAL;
  }
  memset(out, 0, sizeof(*out));
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if ((int64_t)nt > n_bytes / (1 << 20) + 1) nt = (int)(n_bytes / (1 << 20) + 1);
  const size_t n = (size_t)n_bytes;
  auto slice = [&](int t) { return n * (size_t)t / (size_t)nt; };
  // 1. newlines per slice
  std::vector<size_t> cnt(nt + 1, 0);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const uint8_t* p = fastq + slice(t);
        const uint8_t* e = fastq + slice(t + 1);
        size_t c = 0;
        while (p < e) {
          const void* q = memchr(p, '\n', (size_t)(e - p));
          if (!q) break;
          ++c;
          p = static_cast<const uint8_t*>(q) + 1;
        }
        cnt[t + 1] = c;
      });
    for (auto& x : th) x.join();
  }
  for (int t = 0; t < nt; ++t) cnt[t + 1] += cnt[t];
  size_t n_nl = cnt[nt];
  const bool tail_line = n > 0 && fastq[n - 1] != '\n';  // last line without a trailing newline
  size_t n_lines = n_nl + (tail_line ? 1 : 0);
  // line_end[i] = offset of the '\n' (or n for the unterminated tail); line i starts at line_end[i-1] + 1
  std::vector<int64_t> line_end(n_lines + 1);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const uint8_t* p = fastq + slice(t);
        const uint8_t* e = fastq + slice(t + 1);
        size_t i = cnt[t];
        while (p < e) {
          const void* q = memchr(p, '\n', (size_t)(e - p));
          if (!q) break;
          line_end[i++] = static_cast<const uint8_t*>(q) - fastq;
          p = static_cast<const uint8_t*>(q) + 1;
        }
      });
    for (auto& x : th) x.join();
  }
  if (tail_line) line_end[n_nl] = (int64_t)n;
  auto lstart = [&](size_t i) -> int64_t { return i == 0 ? 0 : line_end[i - 1] + 1; };
  auto lend = [&](size_t i) -> int64_t {  // exclusive, '\r' stripped
    int64_t e = line_end[i];
    if (e > lstart(i) && fastq[e - 1] == '\r') --e;
    return e;
  };
  while (n_lines && lend(n_lines - 1) == lstart(n_lines - 1)) --n_lines;  // trailing empty lines
  if (n_lines % 4 != 0) {
    set_error("FASTQ has %zu lines, not a multiple of 4", n_lines);
    return DCB200_EINVAL;
  }
  const size_t R = n_lines / 4;
  out->n_records = (int64_t)R;
  if (R == 0) return DCB200_OK;
  auto alloc = [&](size_t bytes) { return malloc(bytes ? bytes : 1); };
  out->name_off = static_cast<int64_t*>(alloc(R * 8));
  out->name_len = static_cast<int32_t*>(alloc(R * 4));
  out->head_len = static_cast<int32_t*>(alloc(R * 4));
  out->seq_off = static_cast<int64_t*>(alloc(R * 8));
  out->seq_len = static_cast<int32_t*>(alloc(R * 4));
  out->qual_off = static_cast<int64_t*>(alloc(R * 8));
  out->qual_len = static_cast<int32_t*>(alloc(R * 4));
  auto release = [&] {
    free(out->name_off); free(out->name_len); free(out->head_len); free(out->seq_off); free(out->seq_len);
    free(out->qual_off); free(out->qual_len);
    memset(out, 0, sizeof(*out));
  };
  if (!out->name_off || !out->name_len || !out->head_len || !out->seq_off || !out->seq_len || !out->qual_off || !out->qual_len) {
    release();
    set_error("dcb200_index_fastq: out of memory for %zu records", R);
    return DCB200_ENOMEM;
  }
  // 3. records (first error by record number wins)
  std::atomic<long long> bad_rec(-1);
  std::atomic<int> bad_kind(0);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const size_t r0 = R * (size_t)t / (size_t)nt, r1 = R * (size_t)(t + 1) / (size_t)nt;
        for (size_t r = r0; r < r1; ++r) {
          const int64_t hs = lstart(4 * r), he = lend(4 * r);
          const int64_t ss = lstart(4 * r + 1), se = lend(4 * r + 1);
          const int64_t ps = lstart(4 * r + 2);
          const int64_t qs = lstart(4 * r + 3), qe = lend(4 * r + 3);
          int kind = 0;
          if (he <= hs || fastq[hs] != '@') kind = 1;
          else if (lend(4 * r + 2) <= ps || fastq[ps] != '+') kind = 2;
          else if (se - ss != qe - qs) kind = 3;
          else if (se == ss) kind = 4;
          if (kind) {
            long long cur = bad_rec.load();
            while ((cur < 0 || (long long)r < cur) && !bad_rec.compare_exchange_weak(cur, (long long)r)) {}
            if (bad_rec.load() == (long long)r) bad_kind.store(kind);
            continue;
          }
          out->name_off[r] = hs + 1;
          out->head_len[r] = (int32_t)(he - hs - 1);
          int64_t b = hs + 1;
          while (b < he && fastq[b] != ' ' && fastq[b] != '\t') ++b;
          out->name_len[r] = (int32_t)(b - hs - 1);
          out->seq_off[r] = ss;
          out->seq_len[r] = (int32_t)(se - ss);
          out->qual_off[r] = qs;
          out->qual_len[r] = (int32_t)(qe - qs);
        }
      });
    for (auto& x : th) x.join();
  }
  if (bad_rec.load() >= 0) {
    const long long r = bad_rec.load();
    switch (bad_kind.load()) {
      case 1: set_error("FASTQ record does not start with '@' (record %lld)", r); break;
      case 2: set_error("FASTQ separator line does not start with '+' (record %lld)", r); break;
      case 3: set_error("record %lld: sequence and quality lengths differ", r); break;
      default: set_error("empty sequence in FASTQ (record %lld)", r); break;
    }
    release();
    return DCB200_EINVAL;
  }
  return DCB200_OK;
}
