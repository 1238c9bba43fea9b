/* Oracle (TEST INFRASTRUCTURE / reported CPU baseline -- never the product path).
 *
 * Socket-free, HDF5-free C restatement of the receive loop of the reference's
 * native receiver, DataReadout/ReadoutControls/lib/PacketMaster.c:286-397:
 *   - word assembly   packet = ntohl(hi[j])<<32 | ntohl(lo[j])        (:306)
 *   - adr = hi>>24 ; adr==255 closes the second (:329-368), "Corrupted EOS" if
 *     the word is not all ones (:331)
 *   - adr <  NPIXELS_PER_ROACH : photons[r][adr][plist]=packet ; if plist <
 *     MAX_EVENTS_PER_SEC-1 { ++plist ; ++photon_counts[sec][r*NPIX+adr] } (:371-381)
 *   - adr >= NPIXELS_PER_ROACH : "non-pixel" warning, dropped             (:382-386)
 *   - words after sec == exptime are ignored                              (:327)
 * The reference file itself cannot be compiled here (needs hdf5.h / h5cc), so
 * this is a "port" baseline, not a "reference" one.
 *
 * Build:  make -C oracle     ->  oracle/_build/libpm_core.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <arpa/inet.h>

#define BUFSIZE_INTS 8192

typedef struct {
    int64_t n_eos, n_corrupt_eos, n_nonpixel, n_ignored, n_stored;
} pm_stats;

/* One roach stream of flat host-order u64 words.
 * counts:  [exptime][n_pix_total] int32, incremented in place.
 * lists:   optional [exptime][n_pix_total][max_events] u64 (may be NULL).
 * hist:    optional [n_pix_total][n_bins] u32 of bin_lut[field] (may be NULL);
 *          field_shift selects the 12-bit field (44 peak, 32 p1, 20 base).   */
int pm_core_words(const uint64_t *words, int64_t n, int roach, int npix_per_roach,
                  int n_pix_total, int exptime, int max_events, int32_t *counts,
                  uint64_t *lists, uint32_t *hist, int n_bins, int field_shift,
                  const uint16_t *bin_lut, pm_stats *st)
{
    int sec = 0;
    int32_t *plist = (int32_t *)calloc((size_t)npix_per_roach, sizeof(int32_t));
    if (!plist) return -1;
    for (int64_t j = 0; j < n; ++j) {
        uint64_t packet = words[j];
        uint32_t adr = (uint32_t)(packet >> 56);
        if (sec < exptime) {
            if (adr == 255) {
                if (packet != (uint64_t)(-1)) st->n_corrupt_eos++;
                st->n_eos++;
                memset(plist, 0, (size_t)npix_per_roach * sizeof(int32_t));
                sec++;
            } else if ((int)adr < npix_per_roach) {
                int apix = roach * npix_per_roach + (int)adr;
                int idx = plist[adr];
                if (lists)
                    lists[((size_t)sec * n_pix_total + apix) * max_events + idx] = packet;
                if (idx < max_events - 1) {
                    plist[adr] = idx + 1;
                    counts[(size_t)sec * n_pix_total + apix]++;
                    st->n_stored++;
                }
                if (hist) {
                    uint32_t f = (uint32_t)(packet >> field_shift) & 0xFFFu;
                    uint32_t b = bin_lut ? bin_lut[f] : f;
                    if ((int)b < n_bins) hist[(size_t)apix * n_bins + b]++;
                }
            } else {
                st->n_nonpixel++;
            }
        } else {
            st->n_ignored++;
        }
    }
    free(plist);
    return 0;
}

/* Same, from the wire: n_bundles x (8192 BE u32 low block, 8192 BE u32 high block). */
int pm_core_wire(const uint32_t *wire, int64_t n_bundles, int roach, int npix_per_roach,
                 int n_pix_total, int exptime, int max_events, int32_t *counts,
                 uint32_t *hist, int n_bins, int field_shift, const uint16_t *bin_lut,
                 pm_stats *st)
{
    int sec = 0;
    int32_t *plist = (int32_t *)calloc((size_t)npix_per_roach, sizeof(int32_t));
    if (!plist) return -1;
    for (int64_t b = 0; b < n_bundles; ++b) {
        const uint32_t *low_order_block = wire + (size_t)b * 2 * BUFSIZE_INTS;
        const uint32_t *high_order_block = low_order_block + BUFSIZE_INTS;
        for (int j = 0; j < BUFSIZE_INTS; ++j) {
            uint32_t h = ntohl(high_order_block[j]);
            uint64_t packet = (((uint64_t)h) << 32) | (uint64_t)ntohl(low_order_block[j]);
            uint32_t adr = h >> 24;
            if (sec < exptime) {
                if (adr == 255) {
                    if (packet != (uint64_t)(-1)) st->n_corrupt_eos++;
                    st->n_eos++;
                    memset(plist, 0, (size_t)npix_per_roach * sizeof(int32_t));
                    sec++;
                } else if ((int)adr < npix_per_roach) {
                    int apix = roach * npix_per_roach + (int)adr;
                    int idx = plist[adr];
                    if (idx < max_events - 1) {
                        plist[adr] = idx + 1;
                        counts[(size_t)sec * n_pix_total + apix]++;
                        st->n_stored++;
                    }
                    if (hist) {
                        uint32_t f = (uint32_t)(packet >> field_shift) & 0xFFFu;
                        uint32_t bb = bin_lut ? bin_lut[f] : f;
                        if ((int)bb < n_bins) hist[(size_t)apix * n_bins + bb]++;
                    }
                } else {
                    st->n_nonpixel++;
                }
            } else {
                st->n_ignored++;
            }
        }
    }
    free(plist);
    return 0;
}
