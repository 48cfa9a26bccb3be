// Batched PDB writer (host side, multi-threaded): result materialisation for create_side_chain_ensemble.
//
// The reference formats every atom line of every conformer with Python string formatting
// (libs/libstructure.py:537-545, 916-931; libs/libpdb.py:178-195) - tens of milliseconds per conformer, i.e.
// far slower than growing it on the GPU.  Here the constant part of every line is rendered once (Python,
// `Structure.get_PDB`) into a fixed-width template; per conformer only the three %8.3f coordinate fields are
// filled in.  Rounding reproduces `Structure.coords = x` (libstructure.py:157-160): float32 x*1000 -> rint ->
// /1000, i.e. the integer rintf(x*1000.0f) printed with three decimals.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

static inline void put_fixed3(char *field /* 8 chars */, float x) {
    // right-justified "%8.3f" of rintf(x*1000)/1000
    const float y = rintf(x * 1000.0f);
    const bool neg = signbit(y);                         // "-0.000" for values that round to minus zero, as Python prints
    unsigned long u = (unsigned long)fabsf(y);
    char tmp[24];
    int n = 0;
    // three decimals
    for (int i = 0; i < 3; ++i) { tmp[n++] = (char)('0' + u % 10); u /= 10; }
    tmp[n++] = '.';
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) tmp[n++] = '-';
    if (n > 8) n = 8;                                    // overflowing fields are truncated, like a fixed-width writer would
    memset(field, ' ', 8);
    for (int i = 0; i < n; ++i) field[7 - i] = tmp[i];
}

extern "C" int mcsce_write_pdbs(int n_conf, int n_atoms, const float *h_coords, const uint8_t *h_ok, const int32_t *h_order,
                                const char *line_template, int line_len, int coord_off, const char *dir, int n_threads,
                                int32_t *n_written) {
    // h_coords [n_conf][n_atoms][3] (growth order), h_order[i] = growth index of output line i,
    // line_template: n_atoms lines of line_len bytes (newline included) with blank coordinate fields at
    // [coord_off, coord_off+24).  Writes <dir>/<n>.pdb for every n with h_ok[n] != 0.
    if (n_conf < 0 || n_atoms <= 0 || line_len < coord_off + 24) return -1;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > n_conf) n_threads = n_conf > 0 ? n_conf : 1;
    std::atomic<int> next(0), written(0), failed(0);
    const size_t bytes = (size_t)n_atoms * line_len;
    auto work = [&]() {
        std::vector<char> buf(bytes);
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= n_conf) break;
            if (h_ok && !h_ok[c]) continue;
            memcpy(buf.data(), line_template, bytes);
            const float *xyz = h_coords + (size_t)c * n_atoms * 3;
            for (int i = 0; i < n_atoms; ++i) {
                const float *p = xyz + (size_t)h_order[i] * 3;
                char *f = buf.data() + (size_t)i * line_len + coord_off;
                put_fixed3(f, p[0]); put_fixed3(f + 8, p[1]); put_fixed3(f + 16, p[2]);
            }
            const std::string path = std::string(dir) + "/" + std::to_string(c) + ".pdb";
            FILE *fp = fopen(path.c_str(), "wb");
            if (!fp) { failed.fetch_add(1); continue; }
            const size_t w = fwrite(buf.data(), 1, bytes, fp);
            fclose(fp);
            if (w != bytes) failed.fetch_add(1); else written.fetch_add(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (n_written) *n_written = written.load();
    return failed.load() ? -2 : 0;
}
