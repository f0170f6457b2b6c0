// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Text-VCF implementation of the handful of htslib entry points declared in oracle/stubs/htslib/*.h, so that
// the UNMODIFIED reference sv2nl sources (binary/parser/vcf.hpp, standalone/sv2nl/source/*.cpp) run in this
// container where htslib (samtools/htslib 1.15.1, cmake/htslib.cmake:64) is absent. What it restates of
// htslib's published behaviour, and nothing more:
//   * hts_open/bcf_hdr_read: header = every line starting with '#'; ##contig=<ID=..> lines in order give
//     id[BCF_DT_CTG]; ##INFO=<ID=..,Type=..> lines give the tag types.
//   * bcf_read (VCF text): one record per line; rid from CHROM (a contig missing from the header is appended,
//     as htslib does with a warning); pos = POS - 1; rlen = END - pos if INFO/END is an integer, else
//     strlen(REF); INFO split on ';' into key[=value]; an undeclared tag becomes Type=String.
//   * bcf_get_info_values: -1 tag unknown to the header, -2 declared type != requested type, -3 tag absent
//     from the record; BCF_HT_INT -> int32 values (comma separated), BCF_HT_STR -> the raw value, NUL
//     terminated, *ndst = length + 1, return = length.
// Plain and gzip/bgzip files are both read through zlib's gzFile.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "htslib/tbx.h"
#include "htslib/vcf.h"

struct htsFile {
  gzFile gz = nullptr;
  std::string pending;  // first non-header line, read ahead by bcf_hdr_read
  bool has_pending = false;
};

namespace {

struct HdrImpl {
  std::map<std::string, int> info_type;  // tag -> BCF_HT_*
  std::map<std::string, int> ctg_id;
  std::vector<char*> ctg_names;  // owned
};

struct RecImpl {
  std::vector<std::pair<std::string, std::string>> info;  // key, value ("" for flags)
};

bool read_line(gzFile gz, std::string& line) {
  line.clear();
  char buf[1 << 16];
  while (gzgets(gz, buf, sizeof buf)) {
    line += buf;
    if (!line.empty() && line.back() == '\n') {
      line.pop_back();
      if (!line.empty() && line.back() == '\r') line.pop_back();
      return true;
    }
  }
  return !line.empty();
}

int add_contig(bcf_hdr_t* h, const std::string& name) {
  auto* im = static_cast<HdrImpl*>(h->impl);
  auto it = im->ctg_id.find(name);
  if (it != im->ctg_id.end()) return it->second;
  const int id = (int)im->ctg_names.size();
  im->ctg_names.push_back(strdup(name.c_str()));
  im->ctg_id[name] = id;
  h->id[BCF_DT_CTG] = static_cast<bcf_idpair_t*>(realloc(h->id[BCF_DT_CTG], sizeof(bcf_idpair_t) * (id + 1)));
  for (int i = 0; i <= id; ++i) { h->id[BCF_DT_CTG][i].key = im->ctg_names[i]; h->id[BCF_DT_CTG][i].val = nullptr; }
  h->n[BCF_DT_CTG] = id + 1;
  return id;
}

// value of `key=` inside a "##X=<...>" header line ("" if absent)
std::string header_attr(const std::string& line, const char* key) {
  const std::string pat = std::string(key) + "=";
  size_t lt = line.find('<');
  if (lt == std::string::npos) return "";
  size_t p = lt + 1;
  while (p < line.size()) {
    size_t e = p;
    bool quoted = false;
    while (e < line.size() && (quoted || (line[e] != ',' && line[e] != '>'))) {
      if (line[e] == '"') quoted = !quoted;
      ++e;
    }
    if (line.compare(p, pat.size(), pat) == 0) return line.substr(p + pat.size(), e - p - pat.size());
    p = e + 1;
  }
  return "";
}

int type_code(const std::string& t) {
  if (t == "Integer") return BCF_HT_INT;
  if (t == "Float") return BCF_HT_REAL;
  if (t == "Flag") return BCF_HT_FLAG;
  return BCF_HT_STR;  // String, Character
}

}  // namespace

extern "C" {

htsFile* hts_open(const char* fn, const char* /*mode*/) {
  gzFile gz = gzopen(fn, "rb");
  if (!gz) return nullptr;
  gzbuffer(gz, 1 << 18);
  auto* f = new htsFile();
  f->gz = gz;
  return f;
}

int hts_close(htsFile* fp) {
  if (!fp) return 0;
  if (fp->gz) gzclose(fp->gz);
  delete fp;
  return 0;
}

void hts_itr_destroy(hts_itr_t*) {}

bcf_hdr_t* bcf_hdr_read(htsFile* fp) {
  if (!fp) return nullptr;
  auto* h = static_cast<bcf_hdr_t*>(calloc(1, sizeof(bcf_hdr_t)));
  auto* im = new HdrImpl();
  h->impl = im;
  std::string line;
  while (read_line(fp->gz, line)) {
    if (line.empty()) continue;
    if (line[0] != '#') {
      fp->pending = line;
      fp->has_pending = true;
      break;
    }
    if (line.rfind("##contig=", 0) == 0) {
      const std::string id = header_attr(line, "ID");
      if (!id.empty()) add_contig(h, id);
    } else if (line.rfind("##INFO=", 0) == 0) {
      const std::string id = header_attr(line, "ID");
      if (!id.empty()) im->info_type[id] = type_code(header_attr(line, "Type"));
    } else if (line.rfind("#CHROM", 0) == 0) {
      break;
    }
  }
  return h;
}

void bcf_hdr_destroy(bcf_hdr_t* h) {
  if (!h) return;
  auto* im = static_cast<HdrImpl*>(h->impl);
  if (im) {
    for (char* p : im->ctg_names) free(p);
    delete im;
  }
  free(h->id[BCF_DT_CTG]);
  free(h);
}

bcf1_t* bcf_init(void) {
  auto* v = static_cast<bcf1_t*>(calloc(1, sizeof(bcf1_t)));
  v->impl = new RecImpl();
  return v;
}

void bcf_destroy(bcf1_t* v) {
  if (!v) return;
  delete static_cast<RecImpl*>(v->impl);
  free(v);
}

static int parse_record(const std::string& line, const bcf_hdr_t* h_const, bcf1_t* v) {
  auto* h = const_cast<bcf_hdr_t*>(h_const);  // htslib, too, adds unseen contigs/tags to the header while parsing
  auto* him = static_cast<HdrImpl*>(h->impl);
  auto* rec = static_cast<RecImpl*>(v->impl);
  rec->info.clear();
  std::vector<std::string> col;
  size_t p = 0;
  while (col.size() < 8) {
    size_t e = line.find('\t', p);
    if (e == std::string::npos) { col.push_back(line.substr(p)); p = line.size(); break; }
    col.push_back(line.substr(p, e - p));
    p = e + 1;
  }
  if (col.size() < 8) return -2;
  v->rid = add_contig(h, col[0]);
  v->pos = (hts_pos_t)strtoll(col[1].c_str(), nullptr, 10) - 1;
  v->rlen = (hts_pos_t)col[3].size();
  const std::string& info = col[7];
  if (info != ".") {
    size_t a = 0;
    while (a <= info.size()) {
      size_t e = info.find(';', a);
      if (e == std::string::npos) e = info.size();
      if (e > a) {
        const std::string item = info.substr(a, e - a);
        const size_t eq = item.find('=');
        std::string key = eq == std::string::npos ? item : item.substr(0, eq);
        std::string val = eq == std::string::npos ? std::string() : item.substr(eq + 1);
        if (!him->info_type.count(key)) him->info_type[key] = BCF_HT_STR;  // htslib: dummy String definition
        rec->info.emplace_back(std::move(key), std::move(val));
      }
      a = e + 1;
    }
  }
  for (auto& kv : rec->info)
    if (kv.first == "END" && him->info_type["END"] == BCF_HT_INT && !kv.second.empty() && kv.second != ".") {
      const hts_pos_t end = strtoll(kv.second.c_str(), nullptr, 10);
      if (end > v->pos) v->rlen = end - v->pos;
    }
  return 0;
}

int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v) {
  if (!fp || !h || !v) return -2;
  std::string line;
  for (;;) {
    if (fp->has_pending) {
      line.swap(fp->pending);
      fp->has_pending = false;
    } else if (!read_line(fp->gz, line)) {
      return -1;
    }
    if (line.empty() || line[0] == '#') continue;
    return parse_record(line, h, v);
  }
}

int vcf_parse(kstring_t* s, const bcf_hdr_t* h, bcf1_t* v) {
  if (!s || !s->s) return -2;
  return parse_record(std::string(s->s, s->l), h, v);
}

const char* bcf_seqname_safe(const bcf_hdr_t* hdr, const bcf1_t* rec) {
  if (!hdr || !rec || rec->rid < 0 || rec->rid >= hdr->n[BCF_DT_CTG]) return "(unknown)";
  return hdr->id[BCF_DT_CTG][rec->rid].key;
}

int bcf_hdr_name2id(const bcf_hdr_t* hdr, const char* id) {
  auto* im = static_cast<HdrImpl*>(hdr->impl);
  auto it = im->ctg_id.find(id);
  return it == im->ctg_id.end() ? -1 : it->second;
}

int bcf_get_info_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type) {
  auto* him = static_cast<HdrImpl*>(hdr->impl);
  auto* rec = static_cast<RecImpl*>(line->impl);
  auto t = him->info_type.find(tag);
  if (t == him->info_type.end()) return -1;
  if (t->second != (type & 0xff)) return -2;
  const std::string* val = nullptr;
  for (auto& kv : rec->info)
    if (kv.first == tag) { val = &kv.second; break; }
  if (!val) return -3;
  if ((type & 0xff) == BCF_HT_STR) {
    const int len = (int)val->size();
    if (*ndst < len + 1) {
      *ndst = len + 1;
      *dst = realloc(*dst, (size_t)*ndst);
    }
    memcpy(*dst, val->data(), (size_t)len);
    static_cast<char*>(*dst)[len] = 0;
    return len;
  }
  if ((type & 0xff) == BCF_HT_INT) {
    std::vector<long long> vals;
    size_t a = 0;
    while (a <= val->size()) {
      size_t e = val->find(',', a);
      if (e == std::string::npos) e = val->size();
      const std::string item = val->substr(a, e - a);
      if (!item.empty() && item != ".") vals.push_back(strtoll(item.c_str(), nullptr, 10));
      a = e + 1;
    }
    if (vals.empty()) return 0;
    const size_t width = type == BCF_HT_LONG ? 8 : 4;
    if (*ndst < (int)vals.size()) {
      *ndst = (int)vals.size();
      *dst = realloc(*dst, width * vals.size());
    }
    for (size_t i = 0; i < vals.size(); ++i) {
      if (width == 8) static_cast<int64_t*>(*dst)[i] = vals[i];
      else static_cast<int32_t*>(*dst)[i] = (int32_t)vals[i];
    }
    return (int)vals.size();
  }
  return -2;  // REAL / FLAG are not used by the reference's sv2nl
}

// ---- tabix: declared so that vcf.hpp compiles; sv2nl never queries by region -------------------------
tbx_t* tbx_index_load(const char*) { return nullptr; }
void tbx_destroy(tbx_t*) {}
int tbx_name2id(tbx_t*, const char*) { return -1; }
hts_itr_t* tbx_itr_queryi(tbx_t*, int, hts_pos_t, hts_pos_t) { return nullptr; }
hts_itr_t* tbx_itr_querys(tbx_t*, const char*) { return nullptr; }
int tbx_itr_next(htsFile*, tbx_t*, hts_itr_t*, void*) { return -2; }

}  // extern "C"
