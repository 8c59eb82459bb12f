// TrajectoryWriter.hpp -- trajectory files in the reference's wire format (C++ twin of export.py).
//
// The MATLAB scripts write `fprintf(file, '%d %d %d\n', [x, y, z])` (trotting/quad_as_bip_no_plots.m:438-439,482-509;
// walking/quad_walk_no_plots.m:507-513,563-613): a double that is not integer-valued comes out as %e, an
// integer-valued one as an integer.  AMR_code_DART/Controller.cpp:147-281 reads three numbers per line.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>

namespace ismpc_host {

inline void format_value(char* buf, size_t cap, double v)
{
    if (std::isfinite(v) && v == std::floor(v) && std::fabs(v) < 9007199254740992.0) snprintf(buf, cap, "%lld", (long long)v);
    else snprintf(buf, cap, "%e", v);
}

// rows: n x 3 doubles, row-major.  Returns false if the file cannot be opened.
inline bool write_rows(const std::string& path, const double* rows, size_t n)
{
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    char a[40], b[40], c[40];
    for (size_t i = 0; i < n; ++i) {
        format_value(a, sizeof a, rows[3 * i]); format_value(b, sizeof b, rows[3 * i + 1]); format_value(c, sizeof c, rows[3 * i + 2]);
        fprintf(f, "%s %s %s\n", a, b, c);
    }
    fclose(f);
    return true;
}

// The six files of one run: ComTrajectory_<tag>.txt, ComVelocity_<tag>.txt, foot_{fl,fr,rl,rr}_<tag>.txt.
inline bool write_all(const std::string& dir, const std::string& tag, const double* pos, const double* vel, size_t n_ticks,
                      const double* fl, const double* fr, const double* rl, const double* rr, size_t n_samples)
{
    return write_rows(dir + "/ComTrajectory_" + tag + ".txt", pos, n_ticks) &&
           write_rows(dir + "/ComVelocity_" + tag + ".txt", vel, n_ticks) &&
           write_rows(dir + "/foot_fl_" + tag + ".txt", fl, n_samples) && write_rows(dir + "/foot_fr_" + tag + ".txt", fr, n_samples) &&
           write_rows(dir + "/foot_rl_" + tag + ".txt", rl, n_samples) && write_rows(dir + "/foot_rr_" + tag + ".txt", rr, n_samples);
}

}  // namespace ismpc_host
