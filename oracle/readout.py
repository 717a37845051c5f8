"""CPU restatement (numpy / plain loops) of the post-path electronics -- TEST INFRASTRUCTURE ONLY.

Follows imsim/bleed_trails.py:26-147 (bleed_eimage, bleed_channel, BleedCharge) and imsim/readout.py:153-203
(cte_matrix), :391-401 (CcdReadout.apply_cte), :403-412 (apply_crosstalk), :414-480 (build_amp_images), with
the type promotions numpy >= 2 (NEP 50) applies to the reference's expressions, which decide the rounding:
  * e-images and amp segments are float32 (galsim.ImageF);
  * ``sum(my_channel[y0:y1]) - (y1 - y0) * full_well``: float32 running sum minus a float64 -> float64 excess;
  * ``full_well - imarr[ypix]`` is float32; ``imarr[ypix] += bled`` adds in float64 when ``bled`` is the
    (float64) excess and in float32 otherwise, then rounds to float32;
  * ``cte_matrix @ column`` is a float64 product stored back into float32;
  * crosstalk: float32 amp x float64 coefficient -> float64, summed left to right starting from int 0.
PINNED: tests/golden/readout.npz holds outputs of the reference's own functions (tests/golden/
make_golden_readout.py executes their source); tests/test_oracle_golden.py checks this file against them.
"""
import numpy as np
import scipy.special


def bleed_channel(channel, full_well):
    c = np.array(channel, dtype=np.float32, copy=True)
    n = c.size
    fw32 = np.float32(full_well)
    sat = c > fw32
    runs = []
    y = 0
    while y < n:
        if sat[y]:
            y0 = y
            while y < n and sat[y]:
                y += 1
            runs.append((y0, y))
        else:
            y += 1
    for y0, y1 in runs:
        s = np.float32(0.0)
        for k in range(y0, y1):
            s = np.float32(s + c[k])
        excess = np.float64(s) - np.float64((y1 - y0) * float(full_well))
        c[y0:y1] = fw32

        def bleed(ypix, excess):
            if 0 <= ypix < n:
                room = np.float32(fw32 - c[ypix])  # python float is weak: float32 arithmetic
                if np.float64(room) <= excess:      # min(room, excess) returns room unless excess < room
                    c[ypix] = np.float32(c[ypix] + room)
                    excess = excess - np.float64(room)
                else:
                    c[ypix] = np.float32(np.float64(c[ypix]) + excess)
                    excess = excess - excess
            elif ypix < 0:
                excess = excess - min(float(full_well), excess)
            return excess

        for dy in range(0, max(y0, n - y1)):
            excess = bleed(y0 - dy - 1, excess)
            if excess == 0:
                break
            excess = bleed(y1 + dy, excess)
            if excess == 0:
                break
    return c


def bleed_eimage(eimage, full_well, midline_stop=True):
    e = np.array(eimage, dtype=np.float32, copy=True)
    cols = set(np.where(e > np.float32(full_well))[1])
    ymid = e.shape[0] // 2
    for x in cols:
        if midline_stop:
            e[:ymid, x] = bleed_channel(e[:ymid, x], full_well)
            e[ymid:, x] = bleed_channel(e[ymid:, x], full_well)
        else:
            e[:, x] = bleed_channel(e[:, x], full_well)
    return e


def cte_band(npix, cti, ntransfers=20):
    """band[i, k] = cte_matrix[i, i - k], k = 0..ntransfers (zero where i - k < 0)."""
    band = np.zeros((npix, ntransfers + 1))
    for i in range(1, npix + 1):
        band[i - 1, 0] = (1.0 - cti) ** i
        jmin = max(1, i - ntransfers)
        j = np.arange(jmin, i)
        band[i - 1, (i - j)] = scipy.special.binom(i - 1, i - j) * (1.0 - cti) ** j * cti ** (i - j)
    return band


def apply_cte(amps, pcti, scti, ntransfers=20):
    out = []
    for a in amps:
        a = np.array(a, dtype=np.float32, copy=True)
        ny, nx = a.shape
        if pcti != 0:
            band = cte_band(ny, pcti, ntransfers)
            src = a.astype(np.float64)
            res = np.zeros_like(src)
            for k in range(ntransfers, -1, -1):  # ascending j, like a row of the matrix times the column
                res[k:, :] += band[k:, k, None] * src[:ny - k, :]
            a = res.astype(np.float32)
        if scti != 0:
            band = cte_band(nx, scti, ntransfers)
            src = a.astype(np.float64)
            res = np.zeros_like(src)
            for k in range(ntransfers, -1, -1):
                res[:, k:] += band[None, k:, k] * src[:, :nx - k]
            a = res.astype(np.float32)
        out.append(a)
    return out


def apply_crosstalk(amps, xtalk):
    if xtalk is None:
        return amps
    out = []
    for i, row in enumerate(xtalk):
        acc = 0
        for x, y in zip(amps, row):
            acc = acc + x * np.float64(y)
        out.append(amps[i] + acc)
    return out


def paint_cosmic_rays(image, crs, uniforms, num_crs):
    """imsim/cosmic_rays.py:74-111 (CosmicRays.paint_cr) restated as plain loops: per hit three uniforms (catalogue
    index, x, y), then ``image[y, x] += value`` under numpy's indexing rules.  crs: list of lists of (x0, y0, values)."""
    img = np.array(image, copy=True)
    ny, nx = img.shape
    k = 0
    for _ in range(num_crs):
        index = int(uniforms[k] * len(crs))
        px, py = int(uniforms[k + 1] * nx), int(uniforms[k + 2] * ny)
        k += 3
        cr = crs[index]
        for x0, y0, values in cr:
            for dx, value in enumerate(values):
                y, x = py + y0 - cr[0][1], px + x0 - cr[0][0] + dx
                if y < -ny or y >= ny or x < -nx or x >= nx:
                    continue
                img[y, x] += value
    return img
