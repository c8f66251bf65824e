// disort/disort_formatter.hpp -- fmt::formatter for DisortOptions (tests/test_disort.cpp:55)
#pragma once
#include <fmt/format.h>

#include <sstream>

#include "disort.hpp"

template <>
struct fmt::formatter<disort::DisortOptions> {
  constexpr auto parse(fmt::format_parse_context& ctx) { return ctx.begin(); }

  template <typename FormatContext>
  auto format(const disort::DisortOptions& p, FormatContext& ctx) const {
    auto join = [](std::vector<double> const& v) {
      std::ostringstream s;
      s << "(";
      for (size_t i = 0; i < v.size(); ++i) s << (i ? ", " : "") << v[i];
      s << ")";
      return s.str();
    };
    auto const& d = p.ds();
    return fmt::format_to(ctx.out(),
                          "(header = {}; flags = {}; nwave = {}; ncol = {}; nlyr = {}; nstr = {}; nmom = {}; "
                          "accur = {}; user_tau = {}; user_mu = {}; user_phi = {}; wave_lower = {}; wave_upper = {})",
                          p.header(), p.flags(), p.nwave(), p.ncol(), d.nlyr, d.nstr, d.nmom, p.accur(),
                          join(p.user_tau()), join(p.user_mu()), join(p.user_phi()), join(p.wave_lower()),
                          join(p.wave_upper()));
  }
};
