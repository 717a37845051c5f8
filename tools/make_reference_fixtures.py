"""Pin the parts of the path whose arithmetic lives in GalSim / batoid against the LIVE libraries.

The build container has neither GalSim nor batoid ("parity unpinned" for the TAN-SIP
evaluation, the ray trace, the per-photon silicon model, stamp sizes, phase screens, SecondKick).  On a machine
that has them (any imSim installation), run

    python tools/make_reference_fixtures.py --out tests/golden/live_reference.npz

and commit the file: ``tests/test_live_reference.py`` (to be written against the keys below) then compares the
oracle and the CUDA path with the real thing.  Everything here is deterministic on the reference side -- no
GalSim random numbers are consumed except where stated -- so the comparison is at rounding level:

  trace_*      batoid: random rays through ``batoid.Optic.fromYaml('LSST_r.yaml')`` (and with the camera rotated,
               a detector shifted, a Zernike added to M2), positions / velocities / times / vignetting;
               the flattened telescope (imsim_b200.extract) is stored next to them as a pickle.
  wcs_*        galsim: a TAN-SIP(3) GSFitsWCS evaluated forward (xyToradec) and inverse (radecToxy).
  areas_*      galsim.SiliconSensor.calculate_pixel_areas on a spot image, with and without tree rings: pins
               updatePixelDistortions + pixel polygons (tables, corner ownership, float rounding).
  accum_*      SiliconSensor.accumulate with diffusion_factor = 0 and no wavelengths / angles, brighter-fatter off
               and on: the only random draw left is the rare "random neighbour" fallback, so per-pixel counts are
               (nearly) deterministic given the photon positions: pins insidePixel, neighbour search, nrecalc cadence.
  size_*       GSObject.getGoodImageSize for the proxy profiles of imsim/stamp_utils.py at a few folding thresholds.
  kick2_*      galsim.SecondKick radial profile xValue(r) (to tabulate the sampler) for the default parameters.
  screen_*     power spectrum of one instantiated AtmosphericScreen (normalisation of the synthesis).

UNTESTED in the build container (the imports below fail there); written from the public APIs of GalSim 2.7 and
batoid 0.8.  It only reads the libraries; nothing of imSim's tree is copied.
"""
import argparse
import pickle
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/live_reference.npz")
    ap.add_argument("--n-rays", type=int, default=20000)
    args = ap.parse_args()
    try:
        import batoid
        import galsim
    except ImportError as e:  # pragma: no cover
        print("GalSim / batoid are not importable here (%s): nothing written" % e)
        return 1
    from imsim_b200 import extract

    out = {"galsim_version": galsim.__version__, "batoid_version": getattr(batoid, "__version__", "?")}
    rng = np.random.default_rng(20261017)
    n = args.n_rays

    # ---------------------------------------------------------------- ray trace
    fid = batoid.Optic.fromYaml("LSST_r.yaml")
    variants = {
        "nominal": fid,
        "rotated": fid.withLocallyRotatedOptic("LSSTCamera", batoid.RotZ(np.radians(60.0))),
        "shifted": fid.withLocallyShiftedOptic("Detector", [0.0, 0.0, -1.5e-5]),
        "zernike": fid.withSurface("M2", batoid.Sum([fid["M2"].surface,
                                                     batoid.Zernike([0, 0, 0, 0, 1e-7, 0, 2e-8], R_outer=1.71, R_inner=0.9)])),
    }
    r = np.sqrt(rng.uniform(2.3 ** 2, 4.3 ** 2, n))
    ph = rng.uniform(0, 2 * np.pi, n)
    thx, thy = rng.uniform(-0.031, 0.031, n), rng.uniform(-0.031, 0.031, n)
    wl = rng.uniform(320e-9, 1050e-9, n)
    for tag, tel in variants.items():
        nair = tel.inMedium.getN(wl)
        g = 1.0 / np.sqrt(1 + thx ** 2 + thy ** 2)
        x, y = r * np.cos(ph), r * np.sin(ph)
        z = tel.stopSurface.surface.sag(x, y)
        rv = batoid.RayVector._directInit(x=x.copy(), y=y.copy(), z=z.copy(), vx=thx * g / nair, vy=thy * g / nair,
                                          vz=-g / nair, t=np.zeros(n), wavelength=wl.copy(), flux=np.ones(n),
                                          vignetted=np.zeros(n, dtype=bool), failed=np.zeros(n, dtype=bool),
                                          coordSys=tel.stopSurface.coordSys)
        inp = {k: getattr(rv, k).copy() for k in ("x", "y", "z", "vx", "vy", "vz", "t", "wavelength")}
        tel.trace(rv)
        for k, v in inp.items():
            out["trace_%s_in_%s" % (tag, k)] = v
        for k in ("x", "y", "z", "vx", "vy", "vz", "t", "vignetted", "failed"):
            out["trace_%s_out_%s" % (tag, k)] = np.array(getattr(rv, k))
        out["trace_%s_telescope_pickle" % tag] = np.frombuffer(pickle.dumps(extract.telescope_from_batoid(tel)), dtype=np.uint8)

    # ---------------------------------------------------------------- TAN-SIP
    header = {"CTYPE1": "RA---TAN-SIP", "CTYPE2": "DEC--TAN-SIP", "CRPIX1": 2048.5, "CRPIX2": 2002.5,
              "CRVAL1": 17.1, "CRVAL2": -28.6, "CD1_1": -4.1e-5, "CD1_2": 3.7e-5, "CD2_1": 3.7e-5, "CD2_2": 4.1e-5,
              "A_ORDER": 3, "B_ORDER": 3}
    for i in range(4):
        for j in range(4 - i):
            if i + j >= 2:
                header["A_%d_%d" % (i, j)] = float(rng.normal(0, 3e-8 / 10 ** (i + j - 2)))
                header["B_%d_%d" % (i, j)] = float(rng.normal(0, 3e-8 / 10 ** (i + j - 2)))
    wcs = galsim.GSFitsWCS(header=galsim.FitsHeader(header))
    px, py = rng.uniform(0, 4096, n), rng.uniform(0, 4004, n)
    ra, dec = wcs.xyToradec(px, py, units="rad")
    bx, by = wcs.radecToxy(ra, dec, units="rad")
    out.update(wcs_crpix=np.array(wcs.crpix), wcs_cd=np.array(wcs.cd), wcs_ab=np.array(wcs.ab),
               wcs_center=np.array([wcs.center.ra.rad, wcs.center.dec.rad]), wcs_x=px, wcs_y=py, wcs_ra=ra, wcs_dec=dec,
               wcs_back_x=bx, wcs_back_y=by)

    # ---------------------------------------------------------------- silicon sensor
    tr_func = galsim.SiliconSensor.simple_treerings(0.26, 47.0)
    for name in ("lsst_itl_50_8", "lsst_e2v_50_8", "lsst_itl_50_32"):
        for tr in (False, True):
            kw = dict(treering_func=tr_func, treering_center=galsim.PositionD(-1000.0, 300.0)) if tr else {}
            tag = "%s_%s" % (name, "tr" if tr else "notr")
            sensor = galsim.SiliconSensor(name=name, rng=galsim.BaseDeviate(5), diffusion_factor=0.0, nrecalc=10000, **kw)
            im = galsim.ImageF(33, 33, init_value=0)
            yy, xx = np.mgrid[1:34, 1:34]
            im.array[:, :] = 8.0e4 * np.exp(-0.5 * ((xx - 17.3) ** 2 + (yy - 16.6) ** 2) / 1.5 ** 2)
            out["areas_%s_image" % tag] = im.array.copy()
            out["areas_%s" % tag] = sensor.calculate_pixel_areas(im).array.copy()
            # deterministic accumulate: no diffusion, no depth, photons on a fixed grid with unit flux
            m = 400000
            pa = galsim.PhotonArray(m)
            pa.x = 17.0 + rng.normal(0, 1.4, m)
            pa.y = 17.0 + rng.normal(0, 1.4, m)
            pa.flux = np.ones(m)
            for strength in (0.0, 1.0):
                s2 = galsim.SiliconSensor(name=name, strength=strength, rng=galsim.BaseDeviate(7), diffusion_factor=0.0,
                                          nrecalc=10000, **kw)
                img = galsim.ImageD(33, 33, init_value=0)
                img.setCenter(17, 17)
                added = s2.accumulate(pa, img, orig_center=galsim.PositionI(0, 0))
                out["accum_%s_s%d" % (tag, int(strength))] = img.array.copy()
                out["accum_%s_s%d_added" % (tag, int(strength))] = np.float64(added)
            out["accum_%s_x" % tag] = np.array(pa.x)
            out["accum_%s_y" % tag] = np.array(pa.y)

    # ---------------------------------------------------------------- stamp sizes (imsim/stamp_utils.py proxies)
    sizes = []
    for ft in (5e-3, np.exp(-6.0), np.exp(-8.0), np.exp(-10.0)):
        gsp = galsim.GSParams(folding_threshold=ft)
        fwhm_atm = 0.7 * (622.2 / 500.0) ** -0.3 * 1.2 ** 0.6
        fwhm_sys = np.sqrt(0.25 ** 2 + 0.3 ** 2 + 0.08 ** 2) * 1.2 ** 0.6
        psf = galsim.Convolve(galsim.Kolmogorov(fwhm=fwhm_atm, gsparams=gsp), galsim.Gaussian(fwhm=fwhm_sys, gsparams=gsp))
        sizes.append((ft, psf.getGoodImageSize(0.2), galsim.Kolmogorov(fwhm=fwhm_atm, gsparams=gsp).stepk,
                      galsim.Gaussian(fwhm=fwhm_sys, gsparams=gsp).stepk))
    out["size_star"] = np.array(sizes)
    gal = []
    for nser, hlr in ((1.0, 1.0), (4.0, 1.0), (4.0, 0.3), (2.5, 2.0)):
        g = galsim.Sersic(n=nser, half_light_radius=hlr)
        gal.append((nser, hlr, g.getGoodImageSize(0.2), g.stepk))
    out["size_sersic"] = np.array(gal)

    # ---------------------------------------------------------------- second kick, one screen
    sk = galsim.SecondKick(622.2, 0.15, 8.36, 0.61, kcrit=0.2)
    rr = np.concatenate([[0.0], np.logspace(-3, 1.5, 400)])
    out["kick2_r_arcsec"] = rr
    out["kick2_xvalue"] = np.array([sk.xValue(x, 0.0) for x in rr])
    scr = galsim.AtmosphericScreen(screen_size=102.4, screen_scale=0.1, altitude=0.0, r0_500=0.2, L0=25.0,
                                   rng=galsim.BaseDeviate(11))
    scr.instantiate()
    tab = np.array(scr._tab2d.getVals())
    out["screen_table_var"] = np.float64(tab.var())
    f = np.fft.fftfreq(tab.shape[0], 0.1)
    p2 = np.abs(np.fft.fft2(tab)) ** 2
    kk = np.hypot(f[:, None], f[None, :])
    bins = np.logspace(-1.8, 0.6, 25)
    out["screen_psd_k"] = 0.5 * (bins[1:] + bins[:-1])
    out["screen_psd"] = np.array([p2[(kk >= a) & (kk < b)].mean() if ((kk >= a) & (kk < b)).any() else 0.0
                                  for a, b in zip(bins[:-1], bins[1:])])

    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "with", len(out), "arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
