"""CPU restatement (numpy) of stage 1 -- TEST INFRASTRUCTURE ONLY, never imported by the product.

Follows the per-photon sequence GalSim applies when imSim shoots an object
(imsim/stamp.py:727-743 -> galsim drawImage(method='phot'); profile shooters galsim/{sersic,gaussian,box,
knots}.py; galsim/phase_psf.py PhaseScreenPSF._shoot; galsim/phase_screens.py AtmosphericScreen.
_wavefront_gradient via LookupTable2D.gradient; galsim SecondKick.shoot; ChromaticAtmosphere scaling,
imsim/atmPSF.py:298-336).  GalSim is not vendored by the reference and not installed here: the sequence
is restated from its published algorithms -- PARITY UNPINNED against GalSim itself.  What this file pins
is that the CUDA kernel computes exactly this arithmetic for injected uniforms.
"""
import numpy as np

ARCSEC = 206264.80624709636
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
MASK = np.uint64(0xFFFFFFFF)


def philox4(seed, index, stream):
    """Philox4x32-10 as in imsim_b200/csrc/b2_common.cuh (counter = index lo, index hi, stream, golden ratio)."""
    index = np.asarray(index, dtype=np.uint64)
    c = [index & MASK, index >> np.uint64(32), np.full(index.shape, stream, np.uint64),
         np.full(index.shape, 0x9E3779B9, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c


def u01(a, b):
    m = ((a & np.uint64(0xFFFFF)) << np.uint64(32)) | b
    return (m.astype(np.float64) + 0.5) * (1.0 / 4503599627370496.0)


def stage1_uniforms(seed, offset, n, nrand=12):
    idx = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    r = np.empty((nrand, n))
    for q in range(nrand // 2):
        w = philox4(seed, idx, 7 + q)
        r[2 * q] = u01(w[0], w[1])
        r[2 * q + 1] = u01(w[2], w[3])
    return r


def table_at(tab, tmax, u):
    n = tab.shape[-1]
    t = -np.log1p(-u)
    g = np.minimum(t, tmax) * ((n - 1) / tmax)
    i = np.minimum(g.astype(np.int64), n - 2)
    f = g - i
    if tab.ndim == 1:
        a, b = tab[i], tab[i + 1]
    else:
        a, b = tab[0][i], tab[0][i + 1]
    return a + f * (b - a)


def screen_gradient(tab, scale, X, Y):
    """galsim.LookupTable2D(interpolant='linear', edge_mode='wrap').gradient on a periodic grid."""
    npix = tab.shape[0]
    ax, ay = X / scale, Y / scale
    fx0, fy0 = np.floor(ax), np.floor(ay)
    fx, fy = ax - fx0, ay - fy0
    ix0 = np.clip((fx0 - npix * np.floor(fx0 / npix)).astype(np.int64), 0, npix - 1)
    iy0 = np.clip((fy0 - npix * np.floor(fy0 / npix)).astype(np.int64), 0, npix - 1)
    ix1, iy1 = (ix0 + 1) % npix, (iy0 + 1) % npix
    f00, f10 = tab[iy0, ix0].astype(np.float64), tab[iy0, ix1].astype(np.float64)
    f01, f11 = tab[iy1, ix0].astype(np.float64), tab[iy1, ix1].astype(np.float64)
    gx = ((f10 - f00) * (1.0 - fy) + (f11 - f01) * fy) / scale
    gy = ((f01 - f00) * (1.0 - fx) + (f11 - f10) * fx) / scale
    return gx, gy


def stage1_photons(objects, counts, r, cdf=None, cdf_wave=None, psf=None, screens=None, kick=None, luts=None,
                   lut_tmax=14.0):
    """objects: structured array (imsim_b200._abi.OBJECT_DTYPE); counts[j] photons of object j; r: uniforms
    [12, n]; psf: B2Psf.  Returns x, y, flux, wavelength."""
    n = int(np.sum(counts))
    j = np.repeat(np.arange(len(counts)), counts)
    ob = objects[j]
    dx, dy = np.zeros(n), np.zeros(n)
    kind = ob["kind"]
    g = kind == 1
    if g.any():
        rad = np.sqrt(-2.0 * np.log(r[0][g]))
        dx[g], dy[g] = rad * np.cos(2 * np.pi * r[1][g]), rad * np.sin(2 * np.pi * r[1][g])
    g = kind == 2
    if g.any():
        rho = np.zeros(g.sum())
        rows = ob["lut"][g]
        for row in np.unique(rows):
            m = rows == row
            rho[m] = table_at(luts[row], lut_tmax, r[0][g][m])
        dx[g], dy[g] = rho * np.cos(2 * np.pi * r[1][g]), rho * np.sin(2 * np.pi * r[1][g])
    g = kind == 3
    if g.any():
        nk = ob["n_knots"][g]
        k = np.minimum((r[2][g] * nk).astype(np.uint64), (nk - 1).astype(np.uint64))
        dxk, dyk = np.zeros(g.sum()), np.zeros(g.sum())
        seeds = ob["knot_seed"][g]
        for sd in np.unique(seeds):
            m = seeds == sd
            w = philox4(int(sd), k[m], 11)
            rad = np.sqrt(-2.0 * np.log(u01(w[0], w[1]))) * (1.0 / 1.1774100225154747)
            ph = 2 * np.pi * u01(w[2], w[3])
            dxk[m], dyk[m] = rad * np.cos(ph), rad * np.sin(ph)
        dx[g], dy[g] = dxk, dyk
    g = kind == 4
    if g.any():
        dx[g], dy[g] = (r[0][g] - 0.5) * ob["p0"][g], (r[1][g] - 0.5) * ob["p1"][g]
    m = ob["m"]
    px = ob["x"] + (m[:, 0] * dx + m[:, 1] * dy)
    py = ob["y"] + (m[:, 2] * dx + m[:, 3] * dy)
    wave = np.zeros(n)
    if cdf is not None:
        for sd in np.unique(ob["sed"]):
            sel = ob["sed"] == sd
            c, cw = cdf[sd], cdf_wave[sd]
            u = r[3][sel]
            a = np.clip(np.searchsorted(c, u, side="right") - 1, 0, c.size - 2)
            c0, c1 = c[a], c[a + 1]
            f = np.where(c1 > c0, (u - c0) / np.where(c1 > c0, c1 - c0, 1.0), 0.0)
            wave[sel] = cw[a] + f * (cw[a + 1] - cw[a])
    kx, ky = np.zeros(n), np.zeros(n)
    if psf is not None:
        if psf.n_screens > 0:
            ri2, ro2 = psf.r_inner ** 2, psf.r_outer ** 2
            rr = np.sqrt(ri2 + (ro2 - ri2) * r[4])
            pu, pv = rr * np.cos(2 * np.pi * r[5]), rr * np.sin(2 * np.pi * r[5])
            t = psf.t0 + psf.exptime * r[6]
            tx, ty = ob["tanx"], ob["tany"]
            gx, gy = np.zeros(n), np.zeros(n)
            for l in range(psf.n_screens):
                X = pu - psf.vx[l] * t + psf.altitude[l] * tx
                Y = pv - psf.vy[l] * t + psf.altitude[l] * ty
                ax, ay = screen_gradient(screens[l], psf.screen_scale, X, Y)
                gx += ax
                gy += ay
            chrom = (wave / psf.base_wavelength) ** psf.exponent if (cdf is not None and psf.exponent != 0.0) else 1.0
            kx += gx * 1e-9 * ARCSEC * chrom
            ky += gy * 1e-9 * ARCSEC * chrom
        if psf.n_kick > 0:
            sel = r[7] >= psf.kick_delta_prob
            u = (r[7][sel] - psf.kick_delta_prob) / (1.0 - psf.kick_delta_prob)
            th = table_at(kick, psf.kick_tmax, u)
            kx[sel] += th * np.cos(2 * np.pi * r[8][sel])
            ky[sel] += th * np.sin(2 * np.pi * r[8][sel])
        if psf.gauss_sigma > 0:
            rad = psf.gauss_sigma * np.sqrt(-2.0 * np.log(r[9]))
            kx += rad * np.cos(2 * np.pi * r[10])
            ky += rad * np.sin(2 * np.pi * r[10])
        a = psf.arcsec_to_pix
        px = px + (a[0] * kx + a[1] * ky)
        py = py + (a[2] * kx + a[3] * ky)
    return px, py, np.ones(n), wave
