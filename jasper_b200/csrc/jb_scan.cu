// jb_scan.cu -- the bulk scan kernels (design notes in jb_scan.cuh).
#include <algorithm>

#include "jb_scan.cuh"
#include "jb_jobs.cuh"

namespace jb {

// (rc >> 2) | (3 - code) << (2k - 2): the reverse complement of the window after pushing `code` on the right
__device__ __forceinline__ Key rc_push(Key rc, uint32_t code, int top_shift) {
  Key r;
  r.lo = (rc.lo >> 2) | (rc.hi << 62);
  r.hi = rc.hi >> 2;
  const uint64_t c = (uint64_t)(3u - code);
  if (top_shift >= 64) r.hi |= c << (top_shift - 64); else r.lo |= c << top_shift;
  return r;
}

// rare: the count does not fit the slot's field, or the record was displaced beyond reach -- the owner asks the stash
__device__ __noinline__ uint32_t scan_stash_lookup(const TableView& t, const uint8_t* window, int k) {
  const Key m = key_canonical(key_from_chars(window, k, k, t.mask), k);
  return stash_find(t, m, key_posx(t.g, hash_lut(t.lut, t.g.key_bytes, m), m));
}

// rare: a window that is not k chars of ACGTacgt (N / IUPAC / the contig tail) goes through MerDNA's own char loop
#ifdef JB_SCAN_RARE_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
Key scan_rare_window_key(const uint8_t* window, int k, Key mask) {
  return key_canonical(key_from_chars(window, k, k, mask), k);
}

// rare, index scan: the minimizer of such a window's (padded) key, from scratch
__device__ __noinline__ uint32_t scan_rare_min_hash(Key canonical, int k, int m) {
  return window_min_hash(canonical, key_revcomp(canonical, k), k, m, nullptr);
}

template <int kLayout>
__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm)
k_scan(const __grid_constant__ TableView t, const __grid_constant__ ScanArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int k = t.g.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* s_lut = reinterpret_cast<uint64_t*>(smem);
  const int lut_n = kLayout == kScanByIndex ? 0 : t.g.key_bytes * 256;        // (the index scan never hashes with the matrix)
  uint8_t* s_warp = smem + (size_t)lut_n * 8 + (size_t)warp * kScanSmemPerWarp;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_warp + kSeqBufs * kSeqSlice);
  for (int i = threadIdx.x; i < lut_n; i += kScanThreads) s_lut[i] = t.lut[i];
  // index scan: every lane's ring of the m-mer hashes of its current window (odd stride: no bank conflicts)
  uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem + (size_t)lut_n * 8 + (size_t)kScanWarps * kScanSmemPerWarp) +
                     (size_t)threadIdx.x * (size_t)scan_ring_stride(k, t.ix.m);
#if JB_SCAN_TMA
  // one mbarrier per tile buffer, owned by this warp: lane 0 asks the TMA unit for a tile's bases (one 1-D bulk copy, the
  // source is 16-byte aligned: contig offsets are multiples of 32, tile starts of 1024) and every lane waits on the barrier
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_cnt + kCntSlice);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;                                // bit b = parity the next wait on buffer b expects
  int buf = 0;
  auto tile_fetch = [&](uint32_t ti, int into) {     // lane 0 only
    const ScanTile nt = decode_tile(a.tile_first, a.n_units, ti);
    const ScanUnit nu = a.units[nt.unit];
    const uint8_t* src = a.bases + nu.seq_off + nt.start;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_warp + into * kSeqSlice), bar = bar0 + 8 * into;
    // (always the whole slice: a contig's region extends 4096 bytes past its last base, the bytes behind the end are cleared below)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kSeqSlice) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"((uint32_t)kSeqSlice), "r"(bar) : "memory");
  };
#endif
  __syncthreads();                                   // the only block-wide barrier
  unsigned long long steps_total = 0;
  const int top_shift = 2 * k - 2;
  const uint32_t warps_total = gridDim.x * kScanWarps;
  const unsigned long long lim = 1ull << t.g.cnt_bits;
  const uint32_t cmask = cval_mask(t.g);
#if JB_SCAN_TMA
  if (lane == 0 && blockIdx.x * kScanWarps + warp < a.n_tiles) tile_fetch(blockIdx.x * kScanWarps + warp, 0);
#endif

  for (uint32_t ti = blockIdx.x * kScanWarps + warp; ti < a.n_tiles; ti += warps_total) {
    const ScanTile tile = decode_tile(a.tile_first, a.n_units, ti);
    const ScanUnit u = a.units[tile.unit];
    const int64_t n_out = (int64_t)min((uint64_t)kTile, u.len + 1 - tile.start);   // positions produced
    const int n_chars = (int)n_out + k - 1;
    const int64_t avail = (int64_t)u.len - (int64_t)tile.start;                    // real bases from tile.start on
#if JB_SCAN_TMA
    uint8_t* s_seq = s_warp + buf * kSeqSlice;
    __syncwarp();                                                                  // the other buffer's readers (previous tile) are done
    if (lane == 0 && ti + warps_total < a.n_tiles) tile_fetch(ti + warps_total, buf ^ 1);
    {
      const uint32_t bar = bar0 + 8 * buf, par = (phase >> buf) & 1u;
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(par) : "memory");
      phase ^= 1u << buf;
    }
    if (avail < n_chars + 1)                                                       // past the contig end: zeros, as MerDNA sees them
      for (int i = (int)(avail > 0 ? avail : 0) + lane; i < n_chars + 1; i += 32) s_seq[i] = 0;
    __syncwarp();
    buf ^= 1;
#else
    uint8_t* s_seq = s_warp;
    const uint8_t* src = a.bases + u.seq_off + tile.start;
    __syncwarp();                                                                  // previous tile's smem reads are done
    // ---- stage bases: 16-byte loads where the source is aligned, zero-fill past the contig end
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      for (int v = lane; v * 16 < n_chars + 1; v += 32) {
        uint4 w = make_uint4(0, 0, 0, 0);
        if ((int64_t)v * 16 + 16 <= avail) w = *reinterpret_cast<const uint4*>(src + v * 16);
        else {
          uint8_t tmp[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) tmp[q] = ((int64_t)v * 16 + q < avail) ? src[v * 16 + q] : (uint8_t)0;
          w = *reinterpret_cast<uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(s_seq + v * 16) = w;
      }
    } else {
      for (int i = lane; i < n_chars + 1; i += 32) s_seq[i] = (i < avail) ? src[i] : (uint8_t)0;
    }
    __syncwarp();
#endif

    // ---- each lane: 32 consecutive windows starting at r0
    const int r0 = lane * kLanePos;
    const int nr = (int)min((int64_t)kLanePos, n_out - r0);      // may be <= 0
    uint32_t okmask = 0;
    Key fwd = {0, 0}, rc = {0, 0};
    int nv = 0;
    if (nr > 0) {
      for (int j = 0; j < k - 1; ++j) {                           // the first k-1 bases of the first window
        const uint8_t c = s_seq[r0 + j];
        const bool v = base_valid(c);
        const uint32_t code = v ? base_code(c) : 0u;
        nv = v ? nv + 1 : 0;
        fwd = key_push(fwd, code, t.mask);
        rc = rc_push(rc, code, top_shift);
      }
    }
    if (kLayout == kLayoutWide) {
      for (int q0 = 0; q0 < nr; q0 += kGroup) {
        Key keys[kGroup];
        const ulonglong4* sec[kGroup];
        unsigned long long w[kGroup][4];
#pragma unroll
        for (int b = 0; b < kGroup; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            const uint8_t c = s_seq[r0 + q + k - 1];
            const bool v = base_valid(c);
            const uint32_t code = v ? base_code(c) : 0u;
            nv = v ? nv + 1 : 0;
            fwd = key_push(fwd, code, t.mask);
            rc = rc_push(rc, code, top_shift);
            Key m;
            if (nv >= k) { okmask |= 1u << q; m = key_less(rc, fwd) ? rc : fwd; }
            else m = scan_rare_window_key(s_seq + r0 + q, k, t.mask);   // rare: N / IUPAC / contig tail
            keys[b] = m;
            const uint64_t pos = hash_lut(s_lut, t.g.key_bytes, m);
            sec[b] = reinterpret_cast<const ulonglong4*>(unit_ptr(t, home_unit(t.g, pos, m)));
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
          }
        }
#pragma unroll
        for (int b = 0; b < kGroup; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            const Key key = keys[b];
            uint32_t r = 0;
            ++steps_total;
            while (true) {
              const uint32_t c0 = (uint32_t)(w[b][1] >> 32), c1 = (uint32_t)(w[b][3] >> 32);
              if (c0 == 0) { r = 0; break; }
              if (w[b][0] == key.lo && (uint32_t)w[b][1] == (uint32_t)key.hi) { r = c0; break; }
              if (c1 == 0) { r = 0; break; }
              if (w[b][2] == key.lo && (uint32_t)w[b][3] == (uint32_t)key.hi) { r = c1; break; }
              ++sec[b];
              ++steps_total;
              asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                           : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
            }
            s_cnt[lane * 33 + q] = r;
          }
        }
      }
    } else if (kLayout == kScanByIndex) {
      // Scan index (jb_index.cuh): the bucket comes from the window's minimizer, kept up to date as the window rolls (the
      // newest m-mer against the current minimum; a rescan of the k-m+1 m-mers when the minimum has left the window), the
      // bucket's bounds stay in registers while the minimizer does, and the slot is one 16-byte load per window.
      const ScanIndex& ix = t.ix;
      const int m = ix.m, w1 = k - m, wn = w1 + 1;       // wn m-mers per window
      const uint64_t mmask = (1ull << (2 * m)) - 1;
      int rp = 0;                                         // ring slot of the oldest m-mer hash
      uint64_t c_bucket = ~0ull, c_beg = 0;
      uint32_t c_cap = 0;
      for (int q0 = 0; q0 < nr; q0 += kGroupIndex) {
        Key keys[kGroupIndex];
        const ulonglong4* pp[kGroupIndex];                // the pair (32-byte sector) in flight, and the bucket's pairs
        const ulonglong4* pbeg[kGroupIndex];
        uint32_t npair[kGroupIndex], at[kGroupIndex];
        unsigned long long w[kGroupIndex][4];
#pragma unroll
        for (int b = 0; b < kGroupIndex; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            const uint8_t c = s_seq[r0 + q + k - 1];
            const bool v = base_valid(c);
            const uint32_t code = v ? base_code(c) : 0u;
            nv = v ? nv + 1 : 0;
            fwd = key_push(fwd, code, t.mask);
            rc = rc_push(rc, code, top_shift);
            Key mk;
            uint32_t mh;
            if (nv >= k) {
              okmask |= 1u << q;
              mk = key_less(rc, fwd) ? rc : fwd;
              // The minimizer hash = the smallest of the window's wn m-mer hashes.  They live in this lane's ring: a window
              // that follows a valid one replaces the oldest with its newest m-mer; the first valid window after a break
              // (every lane's first window, then only behind an N) fills the ring.  The minimum is taken over the whole ring
              // each time: wn shared-memory loads, but the same instructions in every lane (r2i kept a running minimum and
              // rescanned when it expired -- 8.8 of 32 lanes active on average, 3.5x the instructions).
              if (nv == k) {
                for (int tt = 0; tt < wn; ++tt) s_ring[tt] = mmer_hash_at(fwd, rc, k, m, tt, mmask);
                rp = 0;
              } else {
                s_ring[rp] = mmer_hash_at(fwd, rc, k, m, w1, mmask);
                rp = rp + 1 == wn ? 0 : rp + 1;
              }
              mh = 0xFFFFFFFFu;
              for (int tt = 0; tt < wn; ++tt) mh = min(mh, s_ring[tt]);
            } else {
              mk = scan_rare_window_key(s_seq + r0 + q, k, t.mask);   // rare: N / IUPAC / contig tail
              mh = scan_rare_min_hash(mk, k, m);
            }
            keys[b] = mk;
            const uint64_t bucket = index_bucket(mh, ix.bucket_bits);
            if (bucket != c_bucket) { index_bounds(ix, bucket, &c_beg, &c_cap); c_bucket = bucket; }
            pbeg[b] = reinterpret_cast<const ulonglong4*>(ix.slots + c_beg);     // (bucket offsets and capacities are even: 32-byte aligned)
            npair[b] = c_cap >> 1;
            at[b] = c_cap ? __umulhi(index_slot_hash(mk), c_cap >> 1) : 0u;
            if (c_cap) {
              pp[b] = pbeg[b] + at[b];
              asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(pp[b]));
            }
          }
        }
#pragma unroll
        for (int b = 0; b < kGroupIndex; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            uint32_t cv = 0;
            if (npair[b]) {
              ++steps_total;
              const unsigned long long klo = keys[b].lo;
              const uint32_t khi = (uint32_t)keys[b].hi;
              while (true) {
                // slot = {key.lo, key.hi | count << 32}; the records of a pair fill it front to back
                const uint32_t c0 = (uint32_t)(w[b][1] >> 32), c1 = (uint32_t)(w[b][3] >> 32);
                if (c0 == 0) break;
                if (w[b][0] == klo && (uint32_t)w[b][1] == khi) { cv = c0; break; }
                if (c1 == 0) break;
                if (w[b][2] == klo && (uint32_t)w[b][3] == khi) { cv = c1; break; }
                at[b] = at[b] + 1 == npair[b] ? 0u : at[b] + 1;
                ++steps_total;
                pp[b] = pbeg[b] + at[b];
                asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(pp[b]));
              }
            }
            s_cnt[lane * 33 + q] = cv;
          }
        }
      }
    } else {
      // Compact table: the same loop over four 8-byte slots per sector.  (r2a probed whole 128-byte lines with four lanes
      // each: one line per lookup, but four times the landing registers per probe in flight -- 6.0 ms against 2.2.)
      for (int q0 = 0; q0 < nr; q0 += kGroupCompact) {
        unsigned long long pat[kGroupCompact];
        const ulonglong4* sec[kGroupCompact];
        unsigned long long w[kGroupCompact][4];
#pragma unroll
        for (int b = 0; b < kGroupCompact; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            const uint8_t c = s_seq[r0 + q + k - 1];
            const bool v = base_valid(c);
            const uint32_t code = v ? base_code(c) : 0u;
            nv = v ? nv + 1 : 0;
            fwd = key_push(fwd, code, t.mask);
            rc = rc_push(rc, code, top_shift);
            Key m;
            if (nv >= k) { okmask |= 1u << q; m = key_less(rc, fwd) ? rc : fwd; }
            else m = scan_rare_window_key(s_seq + r0 + q, k, t.mask);   // rare: N / IUPAC / contig tail
            const uint64_t px = key_posx(t.g, hash_lut(s_lut, t.g.key_bytes, m), m);
            pat[b] = compact_pattern(t.g, px, m);
            sec[b] = reinterpret_cast<const ulonglong4*>(unit_ptr(t, home_unit_of_posx(t.g, px)));
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
          }
        }
#pragma unroll
        for (int b = 0; b < kGroupCompact; ++b) {
          const int q = q0 + b;
          if (q < nr) {
            uint32_t cv = 0;
            int hop = 0;
            ++steps_total;
            CompactProbe pr;
            pr.pat = pat[b]; pr.lim = lim; pr.cmask = cmask;
            while (!compact_step(pr, w[b][0], w[b][1], w[b][2], w[b][3], &cv)) {
              if (++hop > kMaxHops) { cv = cmask; break; }          // beyond reach: ask the stash
              ++sec[b];
              ++steps_total;
              asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                           : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
            }
            if (cv == cmask) cv = scan_stash_lookup(t, s_seq + r0 + q, k);
            s_cnt[lane * 33 + q] = cv;
          }
        }
      }
    }
    __syncwarp();
    // ---- stream the counts out coalesced; one validity word per lane
    uint32_t* dst = a.counts + u.out_off + tile.start;
    for (int row = 0; row < 32; ++row) {
      const int i = row * 32 + lane;
      if (i < n_out) dst[i] = s_cnt[row * 33 + lane];
    }
    if (a.okb && r0 < n_out) {                          // 32 validity bytes per lane, two 16-byte stores
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t nib = (okmask >> (4 * q)) & 0xFu;
        w[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      }
      uint4* dst4 = reinterpret_cast<uint4*>(a.okb + u.out_off + tile.start + r0);
      dst4[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst4[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
  if (a.probe_steps) {
    for (int o = 16; o; o >>= 1) steps_total += __shfl_xor_sync(0xffffffffu, steps_total, o);
    if (lane == 0 && steps_total) atomicAdd(a.probe_steps, steps_total);
  }
}

// one warp per 1024-position tile; a lane takes four consecutive positions per step (one 16-byte load of counts, one 4-byte
// load of validity bytes), eight lanes make one 32-bit word of the masks with a single warp reduction (r2d: the per-position
// version issued 59 instructions per 32 positions and was bound by issue slots, not by memory; r2ah: a variant that stages the
// tile's counts in shared memory and gives every lane one whole mask word -- a quarter of the instructions, no reduction --
// measured the same step time, 6.53 ms either way, and was dropped)
__global__ void __launch_bounds__(256)
k_events(const __grid_constant__ EventArgs a) {
  const uint32_t ti = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ti >= a.n_tiles) return;
  const ScanTile tile = decode_tile(a.tile_first, a.n_units, ti);
  const ScanUnit u = a.units[tile.unit];
  const int n = (int)min((uint64_t)kTile, u.len + 1 - tile.start);  // positions of this tile
  const uint64_t f0 = u.out_off + tile.start;                       // flat index of the tile's first position: a multiple of 32
  const uint32_t* cnt_t = a.counts + f0;                            // (16-byte aligned: out_off is a multiple of 32, tile.start of 1024)
  const uint8_t* okb_t = a.okb + f0;
  const int k = a.k;
  const bool head = tile.start == 0;                                // only the first tile of a contig holds positions < k
  const unsigned group = 0xFFu << (lane & 24);                      // the eight lanes that share a word
  const int sh = 4 * (lane & 7);
#pragma unroll 2
  for (int w0 = 0; w0 < n; w0 += 128) {
    const int q0 = w0 + 4 * lane;
    uint32_t c[4] = {0, 0, 0, 0}, prev[4] = {0, 0, 0, 0};
    uint32_t okw = 0;
    if (q0 + 3 < n) {
      const uint4 cv = *reinterpret_cast<const uint4*>(cnt_t + q0);
      c[0] = cv.x; c[1] = cv.y; c[2] = cv.z; c[3] = cv.w;
      okw = *reinterpret_cast<const uint32_t*>(okb_t + q0);
    } else {
#pragma unroll
      for (int x = 0; x < 4; ++x) if (q0 + x < n) { c[x] = cnt_t[q0 + x]; okw |= (uint32_t)okb_t[q0 + x] << (8 * x); }
    }
    // the count k positions back (the 1/50 test); it reaches into the previous tile of the same contig, never before position 0
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int q = q0 + x;
      if (q < n && !(head && q < k)) prev[x] = cnt_t[q - k];
    }
    uint32_t ev = 0, wk = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int q = q0 + x;
      const bool in = q < n;
      const bool ok = in && ((okw >> (8 * x)) & 0xFFu) != 0;
      const bool early = head && q < k;
      const bool low = c[x] < a.solid_thre;
      if (ok && low) wk |= 1u << x;
      if (in && (!ok || early || low || (unsigned long long)c[x] * JB_DIVISOR < (unsigned long long)prev[x])) ev |= 1u << x;
    }
    const uint32_t word = __reduce_or_sync(group, ev << sh), wword = __reduce_or_sync(group, wk << sh);
    if ((lane & 7) == 0 && w0 + 4 * lane < n) {                     // (contig offsets are multiples of 32, not of 1024:
      const uint64_t f = f0 + (uint64_t)(w0 + 4 * lane);            //  a tile can straddle two summary blocks)
      a.events[f >> 5] = word; a.weak[f >> 5] = wword;
      if (word) {
        const uint32_t bit = 1u << ((f >> 10) & 31);
        if (!(atomicOr(&a.events_l1[f >> 15], bit))) atomicOr(&a.events_l2[f >> 20], 1u << ((f >> 15) & 31));   // first bit of its level-1 word
      }
    }
  }
}


// ---- k_recount ----------------------------------------------------------------------------------------------------
// 32 two-bit codes, one per lane (lane 0 = first base), as a 64-bit word with the first base most significant
__device__ __forceinline__ uint64_t spread32(uint32_t v) {
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}
// one round: every lane contributes one base; -> the packed word (same in all lanes) and the validity ballot (bit i = lane i)
__device__ __forceinline__ uint64_t pack_round(uint8_t c, bool in_contig, uint32_t* valid) {
  const bool v = in_contig && base_valid(c);
  const uint32_t code = v ? base_code(c) : 0u;
  *valid = __ballot_sync(0xffffffffu, v);
  const uint32_t p0 = __brev(__ballot_sync(0xffffffffu, code & 1u)), p1 = __brev(__ballot_sync(0xffffffffu, (code >> 1) & 1u));
  return spread32(p0) | (spread32(p1) << 1);
}
// the k bases that start `q` bases into the packed stream a:b:c (32 bases per word, first base most significant), 0 <= q < 32
__device__ __forceinline__ Key key_at(uint64_t a, uint64_t b, uint64_t c, int q, int k, const Key& mask) {
  const int sh = 192 - 2 * q - 2 * k;            // k <= 48, q <= 31: 34 <= sh <= 190
  Key r;
  if (sh >= 128) { const int s = sh - 128; r.lo = a >> s; r.hi = 0; }
  else if (sh >= 64) { const int s = sh - 64; r.lo = s ? (b >> s) | (a << (64 - s)) : b; r.hi = a >> s; }
  else { r.lo = (c >> sh) | (b << (64 - sh)); r.hi = (b >> sh) | (a << (64 - sh)); }
  r.lo &= mask.lo; r.hi &= mask.hi;
  return r;
}
// bits [q, q + k) of the validity stream v0:v1:v2 (bit i of word w = base 32 w + i) all set?
__device__ __forceinline__ bool valid_at(uint32_t v0, uint32_t v1, uint32_t v2, int q, int k) {
  const uint32_t lo = __funnelshift_r(v0, v1, q), mid = __funnelshift_r(v1, v2, q);
  if (k <= 32) { const uint32_t m = k == 32 ? 0xffffffffu : ((1u << k) - 1u); return (lo & m) == m; }
  const uint32_t m = (1u << (k - 32)) - 1u;
  return lo == 0xffffffffu && (mid & m) == m;
}

__device__ __noinline__ Key recount_rare_key(const uint8_t* s, int64_t avail, int k, Key mask) {
  return key_canonical(key_from_chars(s, avail, k, mask), k);
}

struct ProbeInFlight { uint64_t posx; const ulonglong4* sec; unsigned long long w0, w1, w2, w3; };
__device__ __forceinline__ void probe_issue(const TableView& t, const Key& m, ProbeInFlight& p) {
  p.posx = key_posx(t.g, hash_lut(t.lut, t.g.key_bytes, m), m);
  p.sec = reinterpret_cast<const ulonglong4*>(unit_ptr(t, home_unit_of_posx(t.g, p.posx)));
  ldg256(p.sec, p.w0, p.w1, p.w2, p.w3);
}
__device__ __forceinline__ uint32_t probe_finish(const TableView& t, const Key& m, ProbeInFlight& p) {
  if (t.g.layout == kLayoutCompact) {
    const CompactProbe pr = compact_probe_of(t.g, p.posx, m);
    uint32_t cv = 0;
    int hop = 0;
    while (!compact_step(pr, p.w0, p.w1, p.w2, p.w3, &cv)) {
      if (++hop > kMaxHops) { cv = pr.cmask; break; }
      ++p.sec;
      ldg256(p.sec, p.w0, p.w1, p.w2, p.w3);
    }
    return cv == pr.cmask ? stash_find(t, m, p.posx) : cv;
  }
  while (true) {
    const uint32_t c0 = (uint32_t)(p.w1 >> 32), c1 = (uint32_t)(p.w3 >> 32);
    if (c0 == 0) return 0;
    if (p.w0 == m.lo && (uint32_t)p.w1 == (uint32_t)m.hi) return c0;
    if (c1 == 0) return 0;
    if (p.w2 == m.lo && (uint32_t)p.w3 == (uint32_t)m.hi) return c1;
    ++p.sec;
    ldg256(p.sec, p.w0, p.w1, p.w2, p.w3);
  }
}

// Counts of the next pass's sequence, two kernels over the same job list.  k_recount_copy: a job's windows that lie inside
// an untouched stretch keep their counts (streamed, eight independent loads per thread in flight).  k_recount_probe: the
// other windows -- the last k-1 of an untouched stretch, every window of a materialised one -- are cut from a 2-bit packed
// copy of the new bases that the warp builds with ballots, 64 windows per trip, two probes in flight per lane.
// Work split in both: decided on the device from the job count (job_split, jb_jobs.cuh); the launch does not wait for the
// host to learn how many jobs the walkers wrote.
struct RecountJobView { int64_t D, dirty0, dend, cend, O; bool live; };
__device__ __forceinline__ RecountJobView recount_job(const RecountArgs& a, unsigned long long j, int k) {
  RecountJobView v;
  const CopyJob job = a.jobs[j];
  const int64_t off = a.dst_off[job.slot];
  v.live = off >= 0;
  if (!v.live) return v;
  const ScanUnit u = a.units_new[a.slot_unit[job.slot]];
  v.D = off + job.rel;                                            // flat position of the job's first base
  v.cend = (int64_t)u.seq_off + (int64_t)u.len;                   // flat end of the new contig
  v.dend = v.D + job.len;
  v.dirty0 = v.D; v.O = -1;
  if (job.src >= a.in_buf && job.src < a.in_buf + a.flat_bytes) { // an untouched stretch: windows inside it keep their counts
    v.O = job.src - a.in_buf;
    v.dirty0 = v.D + (job.len - k + 1 > 0 ? job.len - k + 1 : 0);
  }
  return v;
}
constexpr int kRecountUnroll = 8;
__global__ void __launch_bounds__(256)
k_recount_copy(const __grid_constant__ RecountArgs a, int k) {
  const unsigned long long n = min(*a.n_jobs, a.jobs_cap);
  const JobSplit sp = job_split(n, a.force_rows != 0, 128u);
  for (unsigned long long j = sp.j0; j < n; j += sp.jstep) {
    const RecountJobView v = recount_job(a, j, k);
    if (!v.live || v.O < 0) continue;
    const int64_t ncopy = v.dirty0 - v.D;
    const uint32_t* __restrict__ co = a.cnt_old + v.O;
    const uint8_t* __restrict__ oo = a.okb_old + v.O;
    uint32_t* __restrict__ cn = a.cnt_new + v.D;
    uint8_t* __restrict__ on = a.okb_new + v.D;
    for (int64_t q0 = sp.tid; q0 < ncopy; q0 += sp.nth * kRecountUnroll) {
      uint32_t c[kRecountUnroll];
      uint8_t o[kRecountUnroll];
#pragma unroll
      for (int x = 0; x < kRecountUnroll; ++x) {
        const int64_t q = q0 + x * sp.nth;
        if (q < ncopy) { c[x] = __ldg(co + q); o[x] = __ldg(oo + q); }
      }
#pragma unroll
      for (int x = 0; x < kRecountUnroll; ++x) {
        const int64_t q = q0 + x * sp.nth;
        if (q < ncopy) { cn[q] = c[x]; on[q] = o[x]; }
      }
    }
  }
}

#ifndef JB_RECOUNT_MINBLOCKS
#define JB_RECOUNT_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, JB_RECOUNT_MINBLOCKS)
k_recount_probe(const __grid_constant__ TableView t, const __grid_constant__ RecountArgs a) {
  const unsigned long long n = min(*a.n_jobs, a.jobs_cap);
  const int k = t.g.k;
  const int lane = threadIdx.x & 31;
  // (rows of at most 8 CTAs: a job rarely has more than a few dozen windows to probe)
  const JobSplit sp = job_split(n, a.force_rows != 0, 8u);
  const int64_t wi = sp.tid >> 5, nw = sp.nth >> 5;               // this warp among the job's warps
  unsigned long long probes = 0;
  for (unsigned long long j = sp.j0; j < n; j += sp.jstep) {
    const RecountJobView jv = recount_job(a, j, k);
    if (!jv.live) continue;
    const int64_t cend = jv.cend, dend = jv.dend;
    for (int64_t c0 = jv.dirty0 + wi * 64; c0 < dend; c0 += nw * 64) {
      // bases [c0, c0 + 128) packed: windows c0 + lane (words 0-2) and c0 + 32 + lane (words 1-3)
      uint64_t w[4];
      uint32_t v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t p = c0 + 32 * r + lane;
        const bool in = p < cend;
        w[r] = pack_round(in ? a.new_bases[p] : (uint8_t)0, in, &v[r]);
      }
      const int64_t pa = c0 + lane, pb = c0 + 32 + lane;
      const bool doa = pa < dend, dob = pb < dend;
      const bool oka = valid_at(v[0], v[1], v[2], lane, k), okb = valid_at(v[1], v[2], v[3], lane, k);
      Key ka = {0, 0}, kb = {0, 0};
      ProbeInFlight fa, fb;
      if (doa) {
        if (oka) { const Key f = key_at(w[0], w[1], w[2], lane, k, t.mask); ka = key_canonical(f, k); }
        else ka = recount_rare_key(a.new_bases + pa, cend - pa, k, t.mask);
        probe_issue(t, ka, fa);
      }
      if (dob) {
        if (okb) { const Key f = key_at(w[1], w[2], w[3], lane, k, t.mask); kb = key_canonical(f, k); }
        else kb = recount_rare_key(a.new_bases + pb, cend - pb, k, t.mask);
        probe_issue(t, kb, fb);
      }
      if (doa) { a.cnt_new[pa] = probe_finish(t, ka, fa); a.okb_new[pa] = oka ? 1 : 0; ++probes; }
      if (dob) { a.cnt_new[pb] = probe_finish(t, kb, fb); a.okb_new[pb] = okb ? 1 : 0; ++probes; }
    }
  }
  if (a.probes) {
    for (int o = 16; o; o >>= 1) probes += __shfl_xor_sync(0xffffffffu, probes, o);
    if (lane == 0 && probes) atomicAdd(a.probes, probes);
  }
}

// the empty tail window of every contig: C('') does not change
__global__ void k_recount_tail(RecountTailArgs a) {
  const uint32_t ui = blockIdx.x * blockDim.x + threadIdx.x;
  if (ui >= a.n_units) return;
  const ScanUnit o = a.units_old[ui], n = a.units_new[ui];
  a.cnt_new[n.out_off + n.len] = a.cnt_old[o.out_off + o.len];
  a.okb_new[n.out_off + n.len] = 0;
}

__global__ void k_compare(const ScanUnit* units, uint32_t n_units, const uint32_t* cnt_a, const uint8_t* okb_a, const uint32_t* cnt_b,
                          const uint8_t* okb_b, unsigned long long* out) {
  unsigned long long bad = 0;
  for (uint32_t ui = blockIdx.y; ui < n_units; ui += gridDim.y) {
    const ScanUnit u = units[ui];
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= u.len; p += (uint64_t)gridDim.x * blockDim.x)
      if (cnt_a[u.out_off + p] != cnt_b[u.out_off + p] || (okb_a[u.out_off + p] != 0) != (okb_b[u.out_off + p] != 0)) ++bad;
  }
  if (bad) atomicAdd(out, bad);
}

int compare_counts(const ScanUnit* d_units, uint32_t n_units, const uint32_t* cnt_a, const uint8_t* okb_a, const uint32_t* cnt_b,
                   const uint8_t* okb_b, unsigned long long* d_scratch, unsigned long long* mismatches, cudaStream_t s) {
  cudaMemsetAsync(d_scratch, 0, 8, s);
  if (n_units) k_compare<<<dim3(64, std::min<uint32_t>(n_units, 1024)), 256, 0, s>>>(d_units, n_units, cnt_a, okb_a, cnt_b, okb_b, d_scratch);
  cudaMemcpyAsync(mismatches, d_scratch, 8, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("compare_counts: ") + cudaGetErrorString(e));
  return JB_OK;
}

int launch_recount(const TableView& view, const RecountArgs& a, const RecountTailArgs& t, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RecountArgs a2 = a;
  a2.force_rows = getenv("JB_RECOUNT_ROWS") != nullptr;           // tests: rows of CTAs whatever the job count
  k_recount_copy<<<(unsigned)sms * 8, 256, 0, s>>>(a2, view.g.k);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("k_recount_copy launch: ") + cudaGetErrorString(e));
  k_recount_probe<<<(unsigned)sms * 2 * JB_RECOUNT_MINBLOCKS, 256, 0, s>>>(view, a2);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("k_recount_probe launch: ") + cudaGetErrorString(e));
  if (t.n_units) {
    k_recount_tail<<<(t.n_units + 127) / 128, 128, 0, s>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("k_recount_tail launch: ") + cudaGetErrorString(e));
  }
  return JB_OK;
}

std::vector<uint32_t> make_tile_prefix(const std::vector<ScanUnit>& units) {
  std::vector<uint32_t> first(units.size() + 1, 0);
  for (size_t u = 0; u < units.size(); ++u) first[u + 1] = first[u] + tiles_of(units[u].len);
  return first;
}

int scan_grid(uint32_t n_tiles) {
  int dev = 0, sms = 148, per_sm = kScanCtasPerSm;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n_tiles + kScanWarps - 1) / kScanWarps, (uint64_t)sms * per_sm));
}

// k_scan needs more than the default 48 KB of dynamic shared memory (hash LUT + 8 warp slices)
int launch_scan(const TableView& view, const ScanArgs& a, cudaStream_t s) {
  const bool by_index = view.ix.bucket_bits && getenv("JB_SCAN_NO_INDEX") == nullptr;
  const size_t smem = by_index ? scan_smem_bytes_index(view.g.k, view.ix.m) : scan_smem_bytes(view.g.key_bytes);
  static thread_local int configured_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    cudaError_t e = cudaFuncSetAttribute(k_scan<kLayoutWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(kMaxKeyBytes));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scan<kLayoutCompact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(kMaxKeyBytes));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scan<kScanByIndex>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes_index(JB_MAX_K, index_mmer_len(JB_MAX_K)));
    if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("cudaFuncSetAttribute(k_scan): ") + cudaGetErrorString(e));
    configured_device = dev;
  }
  if (a.n_tiles == 0) return JB_OK;
  if (by_index) k_scan<kScanByIndex><<<scan_grid(a.n_tiles), kScanThreads, smem, s>>>(view, a);
  else if (view.g.layout == kLayoutCompact) k_scan<kLayoutCompact><<<scan_grid(a.n_tiles), kScanThreads, smem, s>>>(view, a);
  else k_scan<kLayoutWide><<<scan_grid(a.n_tiles), kScanThreads, smem, s>>>(view, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("k_scan launch: ") + cudaGetErrorString(e));
  return JB_OK;
}

int launch_events(const EventArgs& a, cudaStream_t s) {
  if (a.n_tiles == 0) return JB_OK;
  k_events<<<(a.n_tiles + 7) / 8, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(JB_E_CUDA, std::string("k_events launch: ") + cudaGetErrorString(e));
  return JB_OK;
}

}  // namespace jb
