"""Host-side round trip of the neighbour-list layout helpers of adaptive-sph_b200/csrc/lists.cuh (no GPU needed: the
helpers are __host__ __device__ and are compiled here by nvcc into a small host program)."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent(r'''
    #include <cstdio>
    #include <cstdlib>
    #include <vector>
    #include "lists.cuh"
    // Fill one slice (32 columns) the way k_neighbors does, with random segment sizes, then read every entry back with
    // nb_get and check positions never collide and never leave the allocation.
    static uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
    int main() {
      uint32_t seed = 12345u;
      int bad = 0;
      for (int trial = 0; trial < 400; trial++) {
        const bool wide = (trial & 1) != 0;
        const uint32_t i0 = (rnd(seed) % 1000u) * ASPH_PAIR_BLOCK + (rnd(seed) % 8u) * 32u;   // first particle of the slice
        uint32_t cw[32], cf[32], ce[32], chunks = 0;
        for (int l = 0; l < 32; l++) {
          cw[l] = rnd(seed) % 40u; cf[l] = rnd(seed) % 9u; ce[l] = cw[l] + cf[l] + rnd(seed) % 30u;
          const uint32_t c = nb_col_chunks(cw[l], cf[l], ce[l], wide);
          if (c > chunks) chunks = c;
        }
        std::vector<uint16_t> slice(size_t(chunks) * 256u, 0xdead);
        std::vector<uint32_t> far(1100u * ASPH_PAIR_FAR, 0u);
        std::vector<std::vector<uint32_t>> want(32);
        for (uint32_t l = 0; l < 32; l++) {
          const uint32_t i = i0 + l, win0 = nb_win0(i), bias = nb_bias(i);
          for (uint32_t k = 0; k < cw[l]; k++) {           // window rows: contiguous-window slots and far-table slots
            uint32_t j, slot;
            if (k % 5u == 4u) {  // a far-table slot of this tile; its content is a fixed function of (tile, slot)
              slot = ASPH_PAIR_WIN + rnd(seed) % ASPH_PAIR_FAR;
              const uint32_t at = (i / ASPH_PAIR_BLOCK) * ASPH_PAIR_FAR + (slot - ASPH_PAIR_WIN);
              j = 7000000u + at * 3u;
              far[at] = j;
            }
            else { slot = rnd(seed) % ASPH_PAIR_WIN; j = win0 + slot; }
            slice[nb_pos_w(l, k)] = uint16_t(slot * 16u);
            want[l].push_back(j);
          }
          for (uint32_t r = cw[l]; r < nb_pad8(cw[l]); r++) slice[nb_pos_w(l, r)] = uint16_t((i - win0) * 16u);
          for (uint32_t k = 0; k < ce[l] - cw[l]; k++) {   // F then E entries
            const uint32_t kk = k < cf[l] ? k : nb_pad4(cf[l]) + (k - cf[l]);
            const uint32_t j = wide ? 3000000u + rnd(seed) % 4000000u : nb_block0(i) + (rnd(seed) % 60000u) - 30000u;
            const uint32_t pos = nb_pos_fe(l, cw[l], wide, kk);
            if (wide) reinterpret_cast<uint32_t*>(slice.data())[pos] = j; else slice[pos] = uint16_t(j - bias);
            if ((wide ? size_t(pos) * 2u + 1u : size_t(pos)) >= slice.size()) bad++;
            want[l].push_back(j);
          }
        }
        for (uint32_t l = 0; l < 32; l++)
          for (uint32_t k = 0; k < ce[l]; k++)
            if (nb_get(slice.data(), far.data(), wide, i0 + l, k, cw[l], cf[l]) != want[l][k]) bad++;
      }
      // packed counts
      for (uint32_t cwv : {0u, 1u, 17u, 4095u})
        for (uint32_t cfv : {0u, 3u, 20000u})
          for (uint32_t g : {0u, 1u}) {
            const uint32_t c = cwv | (cfv << 12) | (g << 31);
            if (nb_cw(c) != cwv || nb_cf(c) != cfv || nb_cn(c) != cwv + cfv || nb_ghost(c) != (g != 0)) bad++;
          }
      printf("bad=%d\n", bad);
      return bad ? 1 : 0;
    }
''')


def test_list_layout_round_trip(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "layout.cu"
    src.write_text(SRC)
    exe = tmp_path / "layout"
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "adaptive-sph_b200", "csrc"), "-o", str(exe), str(src)],
                          stdout=subprocess.DEVNULL)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr


def test_rows4_schedule_covers_every_row_but_the_own_one():
    """The row schedule of the experimental sweep kernels (solver.cu, for_each_pair<., ., R4>; ASPH_ROWS4=1), restated: with
    the particle's own row last in the W segment, every other row is visited exactly once, in groups of 4, and no visit
    goes beyond the 8-row padding of the column."""
    def visited(cw_real):
        out = []

        def chunk(base, rows_left):
            for half in range(2):
                if half == 1 and rows_left <= 4:
                    break
                out.extend(range(base + 4 * half, base + 4 * half + 4))
        cw = cw_real - 1 if cw_real > 0 else 0
        if cw > 0:
            chunk(0, cw)
        if cw > 8:
            chunk(8, cw - 8)
        k0 = 16
        while cw > 16 and k0 < cw:
            chunk(k0, cw - k0)
            k0 += 8
        return out
    for cw in range(0, 200):
        rows = visited(cw)
        assert len(rows) == len(set(rows))
        assert set(range(max(cw - 1, 0))) <= set(rows)
        assert all(r < (cw + 7) // 8 * 8 for r in rows)
        assert len(rows) == (max(cw - 1, 0) + 3) // 4 * 4
