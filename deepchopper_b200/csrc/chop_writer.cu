// Record assembly + BGZF output of `deepchopper-chop` on host threads (no GPU work in this file).
//
// Reference code replaced: the write loop of src/bin/predict.rs:266-364, the record naming / slicing of
// src/output/split.rs:60-226 ("{id}|s:e|T", "{id}|s:e|I", "{id}|s:e" for --ocq, "@{id}" without description for an
// unchopped read) and the bgzf writer of src/output/writefq.rs.  The per-read decisions (action, adapter / kept intervals)
// come from the GPU (dcb200_smooth_chop*); this side only slices bytes and deflates.
//
// Records are processed in FASTQ order in waves; inside a wave every thread owns a contiguous range of records,
// assembles their text and deflates it into complete BGZF blocks (<= 0xff00 bytes of text each, htslib-compatible), so
// the concatenation of the threads' outputs followed by the 28-byte EOF block is a valid BGZF file whose decompressed
// bytes are exactly the text a sequential writer would have produced.
#include "common.cuh"

#include <zlib.h>

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr size_t kBlockText = 0xff00;
const unsigned char kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct Worker {
  std::string text;                 // pending text (< one block after every record batch)
  std::vector<unsigned char> out;   // finished BGZF blocks
  std::vector<unsigned char> scratch;
  z_stream zs;
  bool z_ok = false;
  int level = 6;
  int64_t records = 0, text_bytes = 0;
  int err = 0;

  bool init(int lvl) {
    level = lvl;
    memset(&zs, 0, sizeof(zs));
    // level 0 = Huffman-only deflate: FASTQ text (near-random bases and qualities) has almost nothing for LZ77 to find,
    // so entropy coding alone gets within 5 % of level 6's size at ~8x its speed
    z_ok = deflateInit2(&zs, level ? level : 6, Z_DEFLATED, -15, 8, level ? Z_DEFAULT_STRATEGY : Z_HUFFMAN_ONLY) == Z_OK;
    scratch.resize(compressBound(kBlockText) + 64);
    return z_ok;
  }
  void done() {
    if (z_ok) deflateEnd(&zs);
    z_ok = false;
  }
  void block(const char* data, size_t n) {
    deflateReset(&zs);
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(data));
    zs.avail_in = (uInt)n;
    zs.next_out = scratch.data();
    zs.avail_out = (uInt)scratch.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) {
      err = 1;
      return;
    }
    size_t clen = scratch.size() - zs.avail_out;
    if (clen + 26 > 65536) {  // incompressible text: store it (cannot happen for <= 0xff00 bytes at level >= 1, but be safe)
      deflateReset(&zs);
      deflateParams(&zs, 0, Z_DEFAULT_STRATEGY);  // stored blocks
      zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(data));
      zs.avail_in = (uInt)n;
      zs.next_out = scratch.data();
      zs.avail_out = (uInt)scratch.size();
      const int rc = deflate(&zs, Z_FINISH);
      deflateParams(&zs, level ? level : 6, level ? Z_DEFAULT_STRATEGY : Z_HUFFMAN_ONLY);
      if (rc != Z_STREAM_END) {
        err = 1;
        return;
      }
      clen = scratch.size() - zs.avail_out;
    }
    const uint16_t bsize = (uint16_t)(clen + 25);  // total block size - 1
    const unsigned char head[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                    (unsigned char)(bsize & 0xff), (unsigned char)(bsize >> 8)};
    out.insert(out.end(), head, head + 18);
    out.insert(out.end(), scratch.data(), scratch.data() + clen);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef*>(data), (uInt)n);
    const uint32_t isize = (uint32_t)n;
    unsigned char tail[8];
    memcpy(tail, &crc, 4);
    memcpy(tail + 4, &isize, 4);
    out.insert(out.end(), tail, tail + 8);
  }
  void drain(bool all) {
    size_t pos = 0;
    while (text.size() - pos >= kBlockText) {
      block(text.data() + pos, kBlockText);
      pos += kBlockText;
    }
    if (all && pos < text.size()) {
      block(text.data() + pos, text.size() - pos);
      pos = text.size();
    }
    text.erase(0, pos);
  }
};

inline void append_int(std::string& s, long long v) {
  char buf[24];
  const int n = snprintf(buf, sizeof(buf), "%lld", v);
  s.append(buf, (size_t)n);
}

// one output record: '@' id [suffix] '\n' seq[s:e] "\n+\n" qual[s:e] '\n'   (slices clamp like Python / Rust `get`)
inline void piece(std::string& t, const uint8_t* id, int id_len, long long s, long long e, char tag, const uint8_t* seq,
                  long long seq_len, const uint8_t* qual, long long qual_len) {
  t.push_back('@');
  t.append(reinterpret_cast<const char*>(id), (size_t)id_len);
  t.push_back('|');
  append_int(t, s);
  t.push_back(':');
  append_int(t, e);
  if (tag) {
    t.push_back('|');
    t.push_back(tag);
  }
  t.push_back('\n');
  const long long s0 = std::min(std::max(s, 0LL), seq_len), e0 = std::min(std::max(e, s0), seq_len);
  t.append(reinterpret_cast<const char*>(seq) + s0, (size_t)(e0 - s0));
  t.append("\n+\n", 3);
  const long long s1 = std::min(std::max(s, 0LL), qual_len), e1 = std::min(std::max(e, s1), qual_len);
  t.append(reinterpret_cast<const char*>(qual) + s1, (size_t)(e1 - s1));
  t.push_back('\n');
}

}  // namespace

extern "C" int dcb200_chop_write_bgzf(const dcb200_fastq_index* ix, int64_t R, const uint8_t* has_pred,
                                      const uint8_t* const* pseq, const int32_t* pseq_len, const uint8_t* action,
                                      const int32_t* n_adapter, const int32_t* adapter_iv, int32_t adapter_stride,
                                      const int32_t* n_keep, const int32_t* keep_iv, int32_t keep_stride, const char* path,
                                      int32_t threads, int32_t level, int64_t* n_records, int64_t* n_text_bytes) {
  return dcb200_chop_write_bgzf_part(ix, R, has_pred, pseq, pseq_len, action, n_adapter, adapter_iv, adapter_stride, n_keep,
                                     keep_iv, keep_stride, path, threads, level, 0, n_records, n_text_bytes);
}

extern "C" int dcb200_chop_write_bgzf_part(const dcb200_fastq_index* ix, int64_t R, const uint8_t* has_pred,
                                           const uint8_t* const* pseq, const int32_t* pseq_len, const uint8_t* action,
                                           const int32_t* n_adapter, const int32_t* adapter_iv, int32_t adapter_stride,
                                           const int32_t* n_keep, const int32_t* keep_iv, int32_t keep_stride,
                                           const char* path, int32_t threads, int32_t level, int32_t flags,
                                           int64_t* n_records, int64_t* n_text_bytes) {
  using dcb::set_error;
  if (!ix || (!ix->fastq && R > 0) || R < 0 || !path || (R > 0 && (!has_pred || !action || !pseq || !pseq_len))) {
    set_error("dcb200_chop_write_bgzf: null argument");
    return DCB200_EINVAL;
  }
  // the interval tables are only read for the actions that need them: check them up front instead of faulting mid-file
  for (int64_t r = 0; r < R; ++r) {
    if (!has_pred[r]) continue;
    const uint8_t a = action[r];
    if (a == DCB200_ACTION_ADAPTERS && (!n_adapter || !adapter_iv || adapter_stride <= 0)) {
      set_error("dcb200_chop_write_bgzf: record %lld needs the adapter intervals (action ADAPTERS) but none were given", (long long)r);
      return DCB200_EINVAL;
    }
    if ((a == DCB200_ACTION_CHOP_T || a == DCB200_ACTION_CHOP_I) && (!n_keep || !keep_iv || keep_stride <= 0)) {
      set_error("dcb200_chop_write_bgzf: record %lld needs the kept intervals (action CHOP) but none were given", (long long)r);
      return DCB200_EINVAL;
    }
  }
  if (level < 0 || level > 9) level = 6;
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min(T, 256));
  FILE* f = fopen(path, (flags & DCB200_WRITE_APPEND) ? "ab" : "wb");
  if (!f) {
    set_error("dcb200_chop_write_bgzf: cannot open '%s' for writing", path);
    return DCB200_EINVAL;
  }
  std::vector<Worker> ws((size_t)T);
  for (auto& w : ws)
    if (!w.init(level)) {
      for (auto& x : ws) x.done();
      fclose(f);
      set_error("dcb200_chop_write_bgzf: deflateInit2 failed");
      return DCB200_ENOMEM;
    }
  const int64_t wave = std::max<int64_t>(4096, (int64_t)T * 2048);  // records per wave (bounds the memory held)
  int64_t total_rec = 0, total_text = 0;
  int rc = DCB200_OK;
  for (int64_t w0 = 0; w0 < R && rc == DCB200_OK; w0 += wave) {
    const int64_t w1 = std::min(R, w0 + wave);
    const int64_t per = (w1 - w0 + T - 1) / T;
    auto work = [&](int t) {
      Worker& w = ws[(size_t)t];
      const int64_t a = std::min(w1, w0 + per * t), b = std::min(w1, a + per);
      for (int64_t r = a; r < b; ++r) {
        if (!has_pred[r]) continue;  // no prediction -> dropped (src/bin/predict.rs:141-144)
        const uint8_t* name = ix->fastq + ix->name_off[r];
        const uint8_t* qual = ix->fastq + ix->qual_off[r];
        const long long qlen = ix->qual_len[r];
        const size_t before = w.text.size();
        switch (action[r]) {
          case DCB200_ACTION_PASSTHROUGH:
            w.text.push_back('@');
            w.text.append(reinterpret_cast<const char*>(name), (size_t)ix->head_len[r]);
            w.text.push_back('\n');
            w.text.append(reinterpret_cast<const char*>(ix->fastq + ix->seq_off[r]), (size_t)ix->seq_len[r]);
            w.text.append("\n+\n", 3);
            w.text.append(reinterpret_cast<const char*>(qual), (size_t)qlen);
            w.text.push_back('\n');
            ++w.records;
            break;
          case DCB200_ACTION_UNCHOPPED:
            w.text.push_back('@');
            w.text.append(reinterpret_cast<const char*>(name), (size_t)ix->name_len[r]);
            w.text.push_back('\n');
            w.text.append(reinterpret_cast<const char*>(pseq[r]), (size_t)pseq_len[r]);
            w.text.append("\n+\n", 3);
            w.text.append(reinterpret_cast<const char*>(qual), (size_t)qlen);
            w.text.push_back('\n');
            ++w.records;
            break;
          case DCB200_ACTION_ADAPTERS:
            for (int i = 0; i < n_adapter[r]; ++i) {
              const int32_t* iv = adapter_iv + ((size_t)r * adapter_stride + i) * 2;
              piece(w.text, name, ix->name_len[r], iv[0], iv[1], 0, pseq[r], pseq_len[r], qual, qlen);
              ++w.records;
            }
            break;
          case DCB200_ACTION_CHOP_T:
          case DCB200_ACTION_CHOP_I: {
            const char tag = action[r] == DCB200_ACTION_CHOP_T ? 'T' : 'I';
            for (int i = 0; i < n_keep[r]; ++i) {
              const int32_t* iv = keep_iv + ((size_t)r * keep_stride + i) * 2;
              piece(w.text, name, ix->name_len[r], iv[0], iv[1], tag, pseq[r], pseq_len[r], qual, qlen);
              ++w.records;
            }
            break;
          }
          default:
            w.err = 2;
            return;
        }
        w.text_bytes += (int64_t)(w.text.size() - before);
        if (w.text.size() >= 4 * kBlockText) w.drain(false);
      }
      // a thread's range ends on a block boundary so that the threads' outputs simply concatenate
      w.drain(true);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (auto& w : ws) {
      if (w.err) {
        set_error(w.err == 2 ? "dcb200_chop_write_bgzf: unknown action code" : "dcb200_chop_write_bgzf: deflate failed");
        rc = DCB200_EINVAL;
      }
      if (rc == DCB200_OK && !w.out.empty() && fwrite(w.out.data(), 1, w.out.size(), f) != w.out.size()) {
        set_error("dcb200_chop_write_bgzf: short write to '%s'", path);
        rc = DCB200_EINVAL;
      }
      w.out.clear();
    }
  }
  for (auto& w : ws) {
    total_rec += w.records;
    total_text += w.text_bytes;
    w.done();
  }
  if (rc == DCB200_OK && !(flags & DCB200_WRITE_NO_EOF) && fwrite(kEof, 1, sizeof(kEof), f) != sizeof(kEof)) {
    set_error("dcb200_chop_write_bgzf: short write to '%s'", path);
    rc = DCB200_EINVAL;
  }
  if (fclose(f) != 0 && rc == DCB200_OK) {
    set_error("dcb200_chop_write_bgzf: close of '%s' failed", path);
    rc = DCB200_EINVAL;
  }
  if (n_records) *n_records = total_rec;
  if (n_text_bytes) *n_text_bytes = total_text;
  return rc;
}
