// sv2nl on the B200 join: same command line as the reference tool (standalone/sv2nl/source/main.cpp:83-165)
//   sv2nl --sv <delly.vcf[.gz]> --non-linear <scannls.vcf[.gz]> [--dis N] [-o out.tsv] [-t N] [-s] [-d] [-h] [-v]
// writes <output>.dup, <output>.inv, <output>.tra (main.cpp:30-32), each starting with the header line.
// -t is accepted and ignored (the GPU join replaces the per-chromosome thread pool); -m (merge) is not
// implemented.
#include <chrono>
#include <exception>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "mapper.hpp"

static void usage() {
  std::puts(
      "Map structural Variation to Non-Linear Transcription (B200 join)\n"
      "Usage: sv2nl [OPTION...]\n"
      "      --sv arg          The file path of segment information from delly\n"
      "      --non-linear arg  The file path of non-linear information from scannls\n"
      "      --dis arg         The distance threshold for trans mapper (default: 1000000)\n"
      "  -o, --output arg      The file path of output (default: output.tsv)\n"
      "  -t, --thread arg      accepted for compatibility, ignored\n"
      "  -s, --short           If running in short read and do not use strand\n"
      "      --device arg      CUDA device (default: 0)\n"
      "  -d, --debug           Print debug info\n"
      "  -h, --help            Print help\n"
      "  -v, --version         Print the current version number");
}

int main(int argc, char** argv) {
  std::string sv_path, nl_path, output = "output.tsv";
  sv2nl::Options opt;
  bool debug = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto value = [&](const char* name) -> std::string {
      if (i + 1 >= argc) { std::fprintf(stderr, "option %s needs a value\n", name); std::exit(2); }
      return argv[++i];
    };
    if (a == "-h" || a == "--help") { usage(); return 0; }
    if (a == "-v" || a == "--version") { std::printf("sv2nl (binary_b200) %s\n", bcu_version()); return 0; }
    if (a == "--sv") sv_path = value("--sv");
    else if (a == "--non-linear") nl_path = value("--non-linear");
    else if (a == "--dis") opt.diff = (std::uint32_t)std::stoul(value("--dis"));
    else if (a == "-o" || a == "--output") output = value("-o");
    else if (a == "-t" || a == "--thread") (void)value("-t");
    else if (a == "-s" || a == "--short") opt.use_strand = false;
    else if (a == "--device") opt.device = std::stoi(value("--device"));
    else if (a == "-d" || a == "--debug") debug = true;
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 2; }
  }
  if (sv_path.empty() || nl_path.empty()) { usage(); return 2; }
  try {
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    auto t0 = now();
    // CUDA start-up (driver initialisation, context, loading the kernels: seconds on a multi-GPU host)
    // runs beside the parsing instead of in front of the first join; and unless the user already restricted
    // the visible devices, only the one device this run uses is exposed to the driver (initialising eight
    // GPUs costs several times as much as initialising one)
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
      setenv("CUDA_VISIBLE_DEVICES", std::to_string(opt.device).c_str(), 1);
      opt.device = 0;
    }
    double warm_s = 0;
    std::thread cuda_warmup([&] {
      auto w0 = now();
      bcu_index* empty = nullptr;
      if (bcu_index_build(opt.device, 0, nullptr, nullptr, nullptr, &empty) == BCU_OK) bcu_index_free(empty);
      warm_s = secs(w0, now());
    });
    // the two files are parsed concurrently (each once; the reference re-parses both per chromosome task)
    sv2nl::VcfTable nl, sv;
    std::exception_ptr sv_error;
    std::thread sv_reader([&] {
      try { sv = sv2nl::read_vcf(sv_path, "delly"); } catch (...) { sv_error = std::current_exception(); }
    });
    try {
      nl = sv2nl::read_vcf(nl_path, "nls");
    } catch (...) {
      sv_reader.join();
      cuda_warmup.join();
      throw;
    }
    sv_reader.join();
    if (sv_error) { cuda_warmup.join(); std::rethrow_exception(sv_error); }
    auto t1 = now();
    if (debug) std::fprintf(stderr, "nl records %zu, sv records %zu, diff %u\n", nl.size(), sv.size(), opt.diff);
    cuda_warmup.join();
    auto t1b = now();
    opt.debug = debug;
    auto res = sv2nl::map_sv2nl(nl, sv, opt);
    auto t2 = now();
    auto write = [&](const char* ext, const sv2nl::Lines& lines) {
      std::ofstream ofs(output + ext, std::ios::binary);
      ofs << sv2nl::HEADER << '\n';
      ofs.write(lines.text.data(), (std::streamsize)lines.text.size());
    };
    write(".dup", res.dup);
    write(".inv", res.inv);
    write(".tra", res.tra);
    auto t3 = now();
    if (debug)
      std::fprintf(stderr, "dup %zu inv %zu tra %zu lines; parse %.3f s (CUDA start-up %.3f s beside it, %.3f s left "
                   "over), map (join + filters) %.3f s, write %.3f s\n",
                   res.dup.size(), res.inv.size(), res.tra.size(), secs(t0, t1), warm_s, secs(t1, t1b), secs(t1b, t2),
                   secs(t2, t3));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "sv2nl: %s\n", e.what());  // the reference's pool swallows this silently (thread_pool.hpp:156-159)
    return 1;
  }
  return 0;
}
