// TEST INFRASTRUCTURE ONLY. Force-included (-include) in front of the UNMODIFIED reference translation units.
//
// The image ships spdlog 1.14 / fmt 10; the reference pins spdlog 1.10 / fmt 8 (library/CMakeLists.txt:38-43).
// fmt >= 9 no longer formats types through their operator<< implicitly, so the reference's own
// spdlog::debug("... {}", record) calls (mapper.cpp:52, helper.hpp:17, interval_tree.hpp:208-215) need the
// explicit opt-in below. It lives here, outside the reference sources, and changes no behaviour.
#pragma once
#include <spdlog/fmt/ostr.h>
#include <spdlog/spdlog.h>

#include <binary/algorithm/interval_tree.hpp>
#include <binary/parser/vcf.hpp>

template <class I>
struct fmt::formatter<binary::algorithm::tree::IntervalNode<I>> : fmt::ostream_formatter {};
template <class K>
struct fmt::formatter<binary::algorithm::tree::BaseInterval<K>> : fmt::ostream_formatter {};
template <class I>
struct fmt::formatter<binary::parser::vcf::BaseVcfRecord<I>> : fmt::ostream_formatter {};
template <class R>
struct fmt::formatter<binary::parser::vcf::BaseVcfInterval<R>> : fmt::ostream_formatter {};
