/*
 * oracle_surface.c -- TEST INFRASTRUCTURE ONLY (see shdom_oracle.h).
 *
 * Plain-C restatement of the surface-reflection part of the SHDOM hot path (paths relative to
 * /root/reference):
 *   src/polarized/shdomsub2.f:1222-1301  SURFACE_BRDF
 *   src/polarized/shdomsub2.f:1303-1357  ROSS_THICK_LI_SPARSE
 *   src/polarized/shdomsub2.f:1359-1520  WAVE_FRESNEL_REFLECTION
 *   src/polarized/shdomsub2.f:1524-1656  DINER_REFLECTION
 *   src/polarized/shdomsub2.f:1661-1699  RPV_REFLECTION
 *   src/ocean_brdf.f:1-332, 529-600      ocean_brdf_sw, morcasiwat, indwat, sunglint, Fresnel, getbound
 *   src/polarized/shdomsub1.f:2597-2669  VARIABLE_BRDF_SURFACE
 *   src/polarized/shdomsub2.f:4756-4790  PLANCK_FUNCTION (UNITS 'T' and radiance units; at3d never
 *                                         uses the band-integrated 'B' units, solver.py:1885)
 * Precisions follow the Fortran declarations: REAL -> float (with float libm), REAL*8 -> double,
 * COMPLEX -> float complex, COMPLEX*16 -> double complex.
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "oracle_internal.h"

#define RF(i, j) reflect[((i) - 1) + 4 * ((j) - 1)]

/* PLANCK_FUNCTION  shdomsub2.f:4756-4790 */
float oracle_planck_function(float temp, int units, const float *waveno, float wavelen)
{
    (void)waveno;
    if (units == 'T') return temp;
    if (temp > 0.0f)
        return 1.1911e8f / (wavelen * wavelen * wavelen * wavelen * wavelen)
               / (expf(1.4388e4f / (wavelen * temp)) - 1);
    return 0.0f;
}

/* RPV_REFLECTION  shdomsub2.f:1661-1699 */
static float rpv_reflection(float rho0, float k, float theta, float mu1, float mu2, float phi)
{
    const float mu_min = 0.03f;
    float x1 = mu1, x2 = mu2, m, f, h, cosphi, sin1, sin2, cosg, tan1, tan2, capg;
    if (x1 < mu_min) x1 = mu_min;
    if (x2 < mu_min) x2 = mu_min;
    m = powf(x1 * x2 * (x1 + x2), k - 1);
    cosphi = cosf(phi);
    sin1 = sqrtf(1.0f - x1 * x1);
    sin2 = sqrtf(1.0f - x2 * x2);
    cosg = x1 * x2 + sin1 * sin2 * cosphi;
    f = (1 - theta * theta) / powf(1 + 2 * theta * cosg + theta * theta, 1.5f);
    tan1 = sin1 / x1;
    tan2 = sin2 / x2;
    capg = sqrtf(fabsf(tan1 * tan1 + tan2 * tan2 - 2 * tan1 * tan2 * cosphi));
    h = 1 + (1 - rho0) / (1 + capg);
    return rho0 * m * f * h;
}

/* ROSS_THICK_LI_SPARSE  shdomsub2.f:1303-1357 */
static void ross_thick_li_sparse(float fiso, float fgeo, float fvol, float hb, float br,
                                 float mudown, float muup, float relaz,
                                 float *reflect_out, float *kgeo_out, float *kvol_out)
{
    const double pi = acos(-1.0);
    double coseta, dsq, cost, mu_d, mu_u, sec_d, sec_u, o, theta_d, theta_u, tan_d, tan_u;
    float kvol, kgeo, r;
    coseta = (double)(mudown * muup + sqrtf(1.0f - mudown * mudown) * sqrtf(1.0f - muup * muup) * cosf(relaz));
    kvol = (float)((((pi / 2.0 - acos(coseta)) * coseta + sqrt(1.0 - coseta * coseta))
                    / (double)(mudown + muup)) - pi / 4.0);
    theta_d = fabs(atan((double)br * sqrt(1.0 - (double)(mudown * mudown)) / (double)mudown));
    theta_u = fabs(atan((double)br * sqrt(1.0 - (double)(muup * muup)) / (double)muup));
    mu_d = cos(theta_d);
    mu_u = cos(theta_u);
    coseta = mu_d * mu_u + sqrt(1.0 - mu_d * mu_d) * sqrt(1.0 - mu_u * mu_u) * (double)cosf(relaz);
    sec_d = 1.0 / mu_d;
    sec_u = 1.0 / mu_u;
    tan_d = sqrt(1.0 - mu_d * mu_d) / mu_d;
    tan_u = sqrt(1.0 - mu_u * mu_u) / mu_u;
    dsq = tan_d * tan_d + tan_u * tan_u - 2 * tan_u * tan_d * (double)cosf(relaz);
    {
        double t = tan_d * tan_u * (double)sinf(relaz);
        cost = (double)hb * sqrt(dsq + t * t) / (sec_d + sec_u);
    }
    if (cost >= 1.0) cost = 1.0;
    if (cost <= -1.0) cost = -1.0;
    o = (acos(cost) - cost * sqrt(1.0 - cost * cost)) * (sec_d + sec_u) / pi;
    kgeo = (float)(o - sec_d - sec_u + 0.5 * (1.0 + coseta) * sec_u * sec_d);
    r = fiso + fvol * kvol + fgeo * kgeo;
    r = fmaxf(r, 0.0f);
    *reflect_out = r; *kgeo_out = kgeo; *kvol_out = kvol;
}

/* WAVE_FRESNEL_REFLECTION  shdomsub2.f:1359-1520 (Stokes dimension <= 3) */
static void wave_fresnel_reflection(float mre, float mim, float windspeed, float mui, float mur,
                                    float phii, float phir, int nstokes, float *reflect)
{
    double sigma2, dmui, dmur, dcosi, dsini, dcosr, dsinr, dsi, dsr;
    double vi1, vi2, vi3, vr1, vr2, vr3, unit1, unit2, unit3, fact1, factor, xi1;
    double ti1, ti2, ti3, tr1, tr2, tr3, pi1, pi2, pi3, pr1, pr2, pr3;
    double pikr, prki, tikr, trki, e1, e2, e3, e4;
    double vp1, vp2, vp3, dmod, rdz2, rdz4, dcoeff, dex, af, af11, af12, af21, af22;
    double p, s1, s2, s3, dcot, t1, t2, shadowi, shadowr, shadow;
    double complex cn1, cn2, cxi2, c1, c2, crper, crpar, cf11, cf12, cf21, cf22;
    double complex c21, c22, ctttp, cttpt, cttpp, ctppt, ctppp, cptpp;
    int i, j;

    cn1 = 1.0;
    cn2 = (double)mre + I * (double)mim;
    sigma2 = fmax(0.0005, 0.0015 + 0.00256 * (double)windspeed);
    dmui = fabs((double)mui);
    dmur = (double)mur;
    if (fabs(dmui - 1.0) < 1e-10) dmui = 0.999999999999;
    if (fabs(dmur - 1.0) < 1e-10) dmur = 0.999999999999;
    dcosi = cos((double)phii);
    dsini = sin((double)phii);
    dcosr = cos((double)phir);
    dsinr = sin((double)phir);
    dsi = sqrt(1.0 - dmui * dmui);
    dsr = sqrt(1.0 - dmur * dmur);
    vi1 = dsi * dcosi; vi2 = dsi * dsini; vi3 = -dmui;
    vr1 = dsr * dcosr; vr2 = dsr * dsinr; vr3 = dmur;
    unit1 = vi1 - vr1; unit2 = vi2 - vr2; unit3 = vi3 - vr3;
    fact1 = unit1 * unit1 + unit2 * unit2 + unit3 * unit3;
    factor = sqrt(1.0 / fact1);
    xi1 = factor * (unit1 * vi1 + unit2 * vi2 + unit3 * vi3);
    cxi2 = csqrt(1.0 - (1.0 - xi1 * xi1) * cn1 * cn1 / (cn2 * cn2));
    c1 = cn1 * xi1;
    c2 = cn2 * cxi2;
    crper = (c1 - c2) / (c1 + c2);
    c1 = cn2 * xi1;
    c2 = cn1 * cxi2;
    crpar = (c1 - c2) / (c1 + c2);
    ti1 = -dmui * dcosi; ti2 = -dmui * dsini; ti3 = -dsi;
    tr1 = dmur * dcosr; tr2 = dmur * dsinr; tr3 = -dsr;
    pi1 = -dsini; pi2 = dcosi; pi3 = 0.0;
    pr1 = -dsinr; pr2 = dcosr; pr3 = 0.0;
    pikr = pi1 * vr1 + pi2 * vr2 + pi3 * vr3;
    prki = pr1 * vi1 + pr2 * vi2 + pr3 * vi3;
    tikr = ti1 * vr1 + ti2 * vr2 + ti3 * vr3;
    trki = tr1 * vi1 + tr2 * vi2 + tr3 * vi3;
    e1 = pikr * prki; e2 = tikr * trki; e3 = tikr * prki; e4 = pikr * trki;
    cf11 = e1 * crper + e2 * crpar;
    cf12 = -e3 * crper + e4 * crpar;
    cf21 = -e4 * crper + e3 * crpar;
    cf22 = e2 * crper + e1 * crpar;
    vp1 = vi2 * vr3 - vi3 * vr2;
    vp2 = vi3 * vr1 - vi1 * vr3;
    vp3 = vi1 * vr2 - vi2 * vr1;
    dmod = (vp1 * vp1 + vp2 * vp2 + vp3 * vp3);
    dmod = dmod * dmod;
    rdz2 = unit3 * unit3;
    rdz4 = rdz2 * rdz2;
    dex = exp(-(unit1 * unit1 + unit2 * unit2) / (2 * sigma2 * rdz2));
    dcoeff = fact1 * fact1 * dex / (4 * dmui * dmur * dmod * rdz4 * 2 * sigma2);
    af = 0.5 * dcoeff;
    af11 = cabs(cf11); af11 = af11 * af11;
    af12 = cabs(cf12); af12 = af12 * af12;
    af21 = cabs(cf21); af21 = af21 * af21;
    af22 = cabs(cf22); af22 = af22 * af22;
    RF(1, 1) = (float)((af11 + af12 + af21 + af22) * af);
    if (nstokes >= 2) {
        RF(1, 2) = (float)((af11 - af12 + af21 - af22) * af);
        RF(2, 1) = (float)((af11 - af22 + af12 - af21) * af);
        RF(2, 2) = (float)((af11 - af12 - af21 + af22) * af);
    }
    c21 = conj(cf21);
    c22 = conj(cf22);
    ctttp = cf11 * conj(cf12);
    cttpt = cf11 * c21;
    cttpp = cf11 * c22;
    ctppt = cf12 * c21;
    ctppp = cf12 * c22;
    cptpp = cf21 * c22;
    if (nstokes >= 3) {
        RF(1, 3) = (float)creal((-ctttp - cptpp) * dcoeff);
        RF(2, 3) = (float)creal((-ctttp + cptpp) * dcoeff);
        RF(3, 1) = (float)creal((-cttpt - ctppp) * dcoeff);
        RF(3, 2) = (float)creal((-cttpt + ctppp) * dcoeff);
        RF(3, 3) = (float)creal((cttpp + ctppt) * dcoeff);
    }
    /* shadowing */
    p = acos(-1.0);
    s1 = sqrt(2 * sigma2 / p);
    s3 = 1.0 / (sqrt(2 * sigma2));
    s2 = s3 * s3;
    dcot = dmui / sqrt(1.0 - dmui * dmui);
    t1 = exp(-s2 * (dcot * dcot));
    t2 = erfc(dcot * s3);
    shadowi = 0.5 * (s1 * t1 / dcot - t2);
    dcot = dmur / sqrt(1.0 - dmur * dmur);
    t1 = exp(-s2 * (dcot * dcot));
    t2 = erfc(dcot * s3);
    shadowr = 0.5 * (s1 * t1 / dcot - t2);
    shadow = 1.0 / (1.0 + shadowi + shadowr);
    for (j = 1; j <= nstokes; j++)
        for (i = 1; i <= nstokes; i++)
            RF(i, j) = (float)(RF(i, j) * shadow);
}

/* DINER_REFLECTION  shdomsub2.f:1524-1656 (Stokes dimension <= 3) */
static void diner_reflection(float a, float k, float b, float zeta, float sigma,
                             float mu1, float mu2, float phi, int nstokes, float *reflect)
{
    float sinth1, sinth2, cosphi, cosscatang, tan1, tan2, capg, hot;
    float gamma, cosgamma, f11, f12, f33, cosbeta, h, sinphi, alpha1, alpha2;
    float cos2alpha1, sin2alpha1, cos2alpha2, sin2alpha2;
    float complex sfcindex, epsilon, d, rp, rs;
    int i, j;
    sfcindex = 1.5f;
    for (j = 1; j <= nstokes; j++) for (i = 1; i <= nstokes; i++) RF(i, j) = 0.0f;
    sinth1 = sqrtf(1.0f - mu1 * mu1);
    sinth2 = sqrtf(1.0f - mu2 * mu2);
    cosphi = cosf(phi);
    cosscatang = -mu1 * mu2 + sinth1 * sinth2 * cosphi;
    cosscatang = fminf(1.0f, fmaxf(-1.0f, cosscatang));
    tan1 = sinth1 / mu1;
    tan2 = sinth2 / mu2;
    capg = sqrtf(fabsf(tan1 * tan1 + tan2 * tan2 + 2 * tan1 * tan2 * cosphi));
    hot = 1 + (1 - a) / (1 + capg);
    RF(1, 1) = a * powf((mu1 + mu2) * mu1 * mu2, k - 1) * expf(b * cosscatang);
    RF(1, 1) = RF(1, 1) * hot;
    if (zeta < 0.0f) return;
    gamma = 0.5f * acosf(-cosscatang);
    cosgamma = cosf(gamma);
    epsilon = sfcindex * sfcindex;
    d = csqrtf(epsilon - 1.0f + cosgamma * cosgamma);
    rp = (epsilon * cosgamma - d) / (epsilon * cosgamma + d);
    rs = (cosgamma - d) / (cosgamma + d);
    {
        float arp = cabsf(rp), ars = cabsf(rs);
        f11 = 0.5f * (arp * arp + ars * ars);
        f12 = 0.5f * (arp * arp - ars * ars);
    }
    f33 = crealf(rp * conjf(rs));
    cosbeta = 0.5f * (mu1 + mu2) / cosgamma;
    if (sigma > 0.0f) {
        const float cb2 = cosbeta * cosbeta;
        h = zeta * expf(-0.5f * (1 / cb2 - 1) / (sigma * sigma)) / (8 * (sigma * sigma) * mu2 * mu1 * (cb2 * cb2));
    } else {
        h = zeta / (8 * mu2 * mu1 * cosbeta);
    }
    RF(1, 1) = RF(1, 1) + h * f11;
    if (nstokes >= 2) {
        sinphi = sinf(phi);
        alpha1 = atanf(sinth2 * sinphi / (mu2 * sinth1 + sinth2 * mu1 * cosphi));
        alpha2 = atanf(sinth1 * sinphi / (sinth2 * mu1 + mu2 * sinth1 * cosphi));
        cos2alpha1 = cosf(2 * alpha1);
        sin2alpha1 = sinf(2 * alpha1);
        cos2alpha2 = cosf(2 * alpha2);
        sin2alpha2 = sinf(2 * alpha2);
        RF(1, 2) = h * f12 * cos2alpha1;
        RF(2, 1) = h * f12 * cos2alpha2;
        RF(2, 2) = h * (f11 * cos2alpha1 * cos2alpha2 + f33 * sin2alpha1 * sin2alpha2);
        if (nstokes >= 3) {
            RF(1, 3) = -h * f12 * sin2alpha1;
            RF(2, 3) = h * (-f11 * sin2alpha1 * cos2alpha2 + f33 * cos2alpha1 * sin2alpha2);
            RF(3, 1) = -h * f12 * sin2alpha2;
            RF(3, 2) = h * (-f11 * cos2alpha1 * sin2alpha2 + f33 * sin2alpha1 * cos2alpha2);
            RF(3, 3) = h * (f11 * sin2alpha1 * sin2alpha2 + f33 * cos2alpha1 * cos2alpha2);
        }
    }
}

/* ---------------- ocean BRDF (src/ocean_brdf.f) ---------------- */
/* getbound  ocean_brdf.f:529-600 (1-based xvals) */
static void getbound(const float *xvals, int ifirst, int ilast, float x, int *ind1, int *ind2)
{
#define XV(i) xvals[(i) - 1]
    int i, imid = ilast / 2 + 1;
    if (XV(ilast) > XV(ifirst)) {
        if (x > XV(imid)) {
            for (i = imid; i <= ilast - 1; i++)
                if (XV(i) <= x && XV(i + 1) >= x) { *ind1 = i; *ind2 = i + 1; return; }
        } else {
            for (i = ifirst; i <= imid; i++)
                if (XV(i) <= x && XV(i + 1) >= x) { *ind1 = i; *ind2 = i + 1; return; }
        }
        if (x < XV(ifirst)) { *ind1 = ifirst; *ind2 = ifirst + 1; }
        else { *ind1 = ilast - 1; *ind2 = ilast; }
    } else {
        if (x < XV(imid)) {
            for (i = imid; i <= ilast - 1; i++)
                if (XV(i) >= x && XV(i + 1) <= x) { *ind1 = i; *ind2 = i + 1; return; }
        } else {
            for (i = ifirst; i <= imid; i++)
                if (XV(i) >= x && XV(i + 1) <= x) { *ind1 = i; *ind2 = i + 1; return; }
        }
        if (x > XV(ifirst)) { *ind1 = ifirst; *ind2 = ifirst + 1; }
        else { *ind1 = ilast - 1; *ind2 = ilast; }
    }
#undef XV
}

/* morcasiwat  ocean_brdf.f:133-234 */
static float morcasiwat(float wl, float c)
{
    static const float tkw[61] = {0.0209f,0.0200f,0.0196f,0.0189f,0.0183f,0.0182f,0.0171f,0.0170f,0.0168f,0.0166f,
        0.0168f,0.0170f,0.0173f,0.0174f,0.0175f,0.0184f,0.0194f,0.0203f,0.0217f,0.0240f,
        0.0271f,0.0320f,0.0384f,0.0445f,0.0490f,0.0505f,0.0518f,0.0543f,0.0568f,0.0615f,
        0.0640f,0.0640f,0.0717f,0.0762f,0.0807f,0.0940f,0.1070f,0.1280f,0.1570f,0.2000f,
        0.2530f,0.2790f,0.2960f,0.3030f,0.3100f,0.3150f,0.3200f,0.3250f,0.3300f,0.3400f,
        0.3500f,0.3700f,0.4050f,0.4180f,0.4300f,0.4400f,0.4500f,0.4700f,0.5000f,0.5500f,0.6500f};
    static const float txc[61] = {0.1100f,0.1110f,0.1125f,0.1135f,0.1126f,0.1104f,0.1078f,0.1065f,0.1041f,0.0996f,
        0.0971f,0.0939f,0.0896f,0.0859f,0.0823f,0.0788f,0.0746f,0.0726f,0.0690f,0.0660f,
        0.0636f,0.0600f,0.0578f,0.0540f,0.0498f,0.0475f,0.0467f,0.0450f,0.0440f,0.0426f,
        0.0410f,0.0400f,0.0390f,0.0375f,0.0360f,0.0340f,0.0330f,0.0328f,0.0325f,0.0330f,
        0.0340f,0.0350f,0.0360f,0.0375f,0.0385f,0.0400f,0.0420f,0.0430f,0.0440f,0.0445f,
        0.0450f,0.0460f,0.0475f,0.0490f,0.0515f,0.0520f,0.0505f,0.0440f,0.0390f,0.0340f,0.0300f};
    static const float te[61] = {0.668f,0.672f,0.680f,0.687f,0.693f,0.701f,0.707f,0.708f,0.707f,0.704f,
        0.701f,0.699f,0.700f,0.703f,0.703f,0.703f,0.703f,0.704f,0.702f,0.700f,
        0.700f,0.695f,0.690f,0.685f,0.680f,0.675f,0.670f,0.665f,0.660f,0.655f,
        0.650f,0.645f,0.640f,0.630f,0.623f,0.615f,0.610f,0.614f,0.618f,0.622f,
        0.626f,0.630f,0.634f,0.638f,0.642f,0.647f,0.653f,0.658f,0.663f,0.667f,
        0.672f,0.677f,0.682f,0.687f,0.695f,0.697f,0.693f,0.665f,0.640f,0.620f,0.600f};
    static const float tbw[61] = {0.0076f,0.0072f,0.0068f,0.0064f,0.0061f,0.0058f,0.0055f,0.0052f,0.0049f,0.0047f,
        0.0045f,0.0043f,0.0041f,0.0039f,0.0037f,0.0036f,0.0034f,0.0033f,0.0031f,0.0030f,
        0.0029f,0.0027f,0.0026f,0.0025f,0.0024f,0.0023f,0.0022f,0.0022f,0.0021f,0.0020f,
        0.0019f,0.0018f,0.0018f,0.0017f,0.0017f,0.0016f,0.0016f,0.0015f,0.0015f,0.0014f,
        0.0014f,0.0013f,0.0013f,0.0012f,0.0012f,0.0011f,0.0011f,0.0010f,0.0010f,0.0010f,
        0.0010f,0.0009f,0.0008f,0.0008f,0.0008f,0.0007f,0.0007f,0.0007f,0.0007f,0.0007f,0.0007f};
    float kw, kd, xc, e, bw, bb, b, bbt, u1, r1, u2, r2, err;
    int iwl;
    if (wl < 0.400f || wl > 0.700f) return 0.000f;
    iwl = 1 + (int)lroundf((wl - 0.400f) / 0.005f);
    kw = tkw[iwl - 1]; xc = txc[iwl - 1]; e = te[iwl - 1]; bw = tbw[iwl - 1];
    if (fabsf(c) < 0.0001f) {
        bb = 0.5f * bw;
        kd = kw;
    } else {
        b = 0.30f * powf(c, 0.62f);
        bbt = 0.002f + 0.02f * (0.5f - 0.25f * log10f(c)) * 0.550f / wl;
        bb = 0.5f * bw + bbt * b;
        kd = kw + xc * powf(c, e);
    }
    u1 = 0.75f;
    r1 = 0.33f * bb / u1 / kd;
    for (;;) {
        u2 = 0.90f * (1.f - r1) / (1.f + 2.25f * r1);
        r2 = 0.33f * bb / u2 / kd;
        err = fabsf((r2 - r1) / r2);
        if (err < 0.0001f) break;
        r1 = r2;
    }
    return r2;
}

/* indwat  ocean_brdf.f:238-318 */
static void indwat(float wl, float xsal, float *nr_out, float *ni_out)
{
    static const float twl[62] = {0.250f,0.275f,0.300f,0.325f,0.345f,0.375f,0.400f,0.425f,0.445f,0.475f,
        0.500f,0.525f,0.550f,0.575f,0.600f,0.625f,0.650f,0.675f,0.700f,0.725f,
        0.750f,0.775f,0.800f,0.825f,0.850f,0.875f,0.900f,0.925f,0.950f,0.975f,
        1.000f,1.200f,1.400f,1.600f,1.800f,2.000f,2.200f,2.400f,2.600f,2.650f,
        2.700f,2.750f,2.800f,2.850f,2.900f,2.950f,3.000f,3.050f,3.100f,3.150f,
        3.200f,3.250f,3.300f,3.350f,3.400f,3.450f,3.500f,3.600f,3.700f,3.800f,3.900f,4.000f};
    static const float tnr[62] = {1.362f,1.354f,1.349f,1.346f,1.343f,1.341f,1.339f,1.338f,1.337f,1.336f,
        1.335f,1.334f,1.333f,1.333f,1.332f,1.332f,1.331f,1.331f,1.331f,1.330f,
        1.330f,1.330f,1.329f,1.329f,1.329f,1.328f,1.328f,1.328f,1.327f,1.327f,
        1.327f,1.324f,1.321f,1.317f,1.312f,1.306f,1.296f,1.279f,1.242f,1.219f,
        1.188f,1.157f,1.142f,1.149f,1.201f,1.292f,1.371f,1.426f,1.467f,1.483f,
        1.478f,1.467f,1.450f,1.432f,1.420f,1.410f,1.400f,1.385f,1.374f,1.364f,1.357f,1.351f};
    static const float tni[62] = {3.35E-08f,2.35E-08f,1.60E-08f,1.08E-08f,6.50E-09f,
        3.50E-09f,1.86E-09f,1.30E-09f,1.02E-09f,9.35E-10f,
        1.00E-09f,1.32E-09f,1.96E-09f,3.60E-09f,1.09E-08f,
        1.39E-08f,1.64E-08f,2.23E-08f,3.35E-08f,9.15E-08f,
        1.56E-07f,1.48E-07f,1.25E-07f,1.82E-07f,2.93E-07f,
        3.91E-07f,4.86E-07f,1.06E-06f,2.93E-06f,3.48E-06f,
        2.89E-06f,9.89E-06f,1.38E-04f,8.55E-05f,1.15E-04f,
        1.10E-03f,2.89E-04f,9.56E-04f,3.17E-03f,6.70E-03f,
        1.90E-02f,5.90E-02f,1.15E-01f,1.85E-01f,2.68E-01f,
        2.98E-01f,2.72E-01f,2.40E-01f,1.92E-01f,1.35E-01f,
        9.24E-02f,6.10E-02f,3.68E-02f,2.61E-02f,1.95E-02f,
        1.32E-02f,9.40E-03f,5.15E-03f,3.60E-03f,3.40E-03f,3.80E-03f,4.60E-03f};
    float nr, ni, xwl, yr, yi;
    const float nrc = 0.006f, nic = 0.000f;
    int i = 2;
    while (!(wl < twl[i - 1]) && i < 62) i++;
    xwl = twl[i - 1] - twl[i - 2];
    yr = tnr[i - 1] - tnr[i - 2];
    yi = tni[i - 1] - tni[i - 2];
    nr = tnr[i - 2] + (wl - twl[i - 2]) * yr / xwl;
    ni = tni[i - 2] + (wl - twl[i - 2]) * yi / xwl;
    if (xsal >= 0.0f) {
        nr = nr + nrc * (xsal / 34.3f);
        ni = ni + nic * (xsal / 34.3f);
    } else {
        nr = nr + nrc;
        ni = ni + nic;
    }
    *nr_out = nr; *ni_out = ni;
}

/* Fresnel  ocean_brdf.f:380-402 */
static float ocean_fresnel(float nr, float ni, float coschi, float sinchi)
{
    float a1, a2, u, v, rr2, rl2, b1, b2, t;
    a1 = fabsf(nr * nr - ni * ni - sinchi * sinchi);
    t = nr * nr - ni * ni - sinchi * sinchi;
    a2 = sqrtf(powf(t, 2.f) + 4 * nr * nr * ni * ni);
    u = sqrtf(0.5f * (a1 + a2));
    v = sqrtf(fmaxf(0.0f, 0.5f * (-a1 + a2)));
    rr2 = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    b1 = (nr * nr - ni * ni) * coschi;
    b2 = 2 * nr * ni * coschi;
    rl2 = ((b1 - u) * (b1 - u) + (b2 - v) * (b2 - v)) / ((b1 + u) * (b1 + u) + (b2 + v) * (b2 + v));
    return (rr2 + rl2) / 2.f;
}

/* sunglint  ocean_brdf.f:322-377 */
static float sunglint(float wspd, float nr, float ni, float azw, float ts, float tv, float fi)
{
    float pi, fac, phw, cs, cv, ss, sv, phi, zx, zy, tantilt, tilt, proba, xe, xn, xe2, xn2;
    float coef, cos2chi, coschi, sinchi, r1, sigmac, sigmau, c21, c03, c40, c04, c22, ct;
    pi = atanf(1.f) * 4.f;
    fac = pi / 180.f;
    phw = azw * fac;
    cs = cosf(ts * fac);
    cv = cosf(tv * fac);
    ss = sinf(ts * fac);
    sv = sinf(tv * fac);
    phi = fi * fac;
    zx = -sv * sinf(phi) / (cs + cv);
    zy = (ss + sv * cosf(phi)) / (cs + cv);
    tantilt = sqrtf(zx * zx + zy * zy);
    tilt = atanf(tantilt);
    sigmac = 0.003f + 0.00192f * wspd;
    sigmau = 0.00316f * wspd;
    c21 = 0.01f - 0.0086f * wspd;
    c03 = 0.04f - 0.033f * wspd;
    c40 = 0.40f;
    c22 = 0.12f;
    c04 = 0.23f;
    xe = (cosf(phw) * zx + sinf(phw) * zy) / sqrtf(sigmac);
    xn = (-sinf(phw) * zx + cosf(phw) * zy) / sqrtf(sigmau);
    xe2 = xe * xe;
    xn2 = xn * xn;
    coef = 1 - c21 / 2.f * (xe2 - 1) * xn - c03 / 6.f * (xn2 - 3) * xn;
    coef = coef + c40 / 24.f * (xe2 * xe2 - 6 * xe2 + 3);
    coef = coef + c04 / 24.f * (xn2 * xn2 - 6 * xn2 + 3);
    coef = coef + c22 / 4.f * (xe2 - 1) * (xn2 - 1);
    proba = coef / 2.f / pi / sqrtf(sigmau) / sqrtf(sigmac) * expf(-(xe2 + xn2) / 2.f);
    cos2chi = cv * cs + sv * ss * cosf(phi);
    if (cos2chi > 1.0f) cos2chi = 0.99999999999f;
    if (cos2chi < -1.0f) cos2chi = -0.99999999999f;
    coschi = sqrtf(0.5f * (1 + cos2chi));
    sinchi = sqrtf(0.5f * (1 - cos2chi));
    r1 = ocean_fresnel(nr, ni, coschi, sinchi);
    ct = cosf(tilt);
    ct = (ct * ct) * (ct * ct);
    return pi * r1 * proba / 4.f / cs / cv / ct;
}

/* ocean_brdf_sw  ocean_brdf.f:1-129 */
static float ocean_brdf_sw(float pws, float xsal, float pcl, float pwl, float xmuo, float xmu,
                           float xphi, float xpaw)
{
    static const float ref[39] = {0.220f,0.220f,0.220f,0.220f,0.220f,0.220f,0.215f,0.210f,0.200f,0.190f,
        0.175f,0.155f,0.130f,0.080f,0.100f,0.105f,0.100f,0.080f,0.045f,0.055f,
        0.065f,0.060f,0.055f,0.040f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,
        0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f};
    static const float angbnd[5] = {0.0f, 45.0f, 60.0f, 75.0f, 85.0f};
    static const float wsbnd[6] = {1.0f, 3.0f, 5.0f, 7.0f, 9.0f, 20.0f};
    /* DATA ((tdsbnd(iang,iws),iws=1,6),iang=1,5): the list runs over iws fastest */
    static const float tdsbnd[5][6] = {
        {0.9787803f,0.9787738f,0.9787626f,0.9787467f,0.9787264f,0.9785573f},
        {0.9706900f,0.9698871f,0.9691746f,0.9685547f,0.9680276f,0.9666586f},
        {0.9479931f,0.9404608f,0.9385692f,0.9381815f,0.9384519f,0.9430056f},
        {0.9690591f,0.9275920f,0.9058769f,0.8951812f,0.8899654f,0.8892645f},
        {0.9980542f,0.9602273f,0.9114283f,0.8713799f,0.8417820f,0.7800314f}};
    static const float tdvbnd[5][6] = {
        {0.9787764f,0.9787535f,0.9787106f,0.9786453f,0.9785548f,0.9775019f},
        {0.9692680f,0.9637051f,0.9564344f,0.9495727f,0.9438773f,0.9288712f},
        {0.9225163f,0.9069787f,0.9044844f,0.9052351f,0.9068328f,0.9153687f},
        {0.8048478f,0.8479503f,0.8678726f,0.8797889f,0.8878716f,0.9091171f},
        {0.7294627f,0.8137348f,0.8453338f,0.8629867f,0.8745421f,0.9036854f}};
    float pi, fac, paw, phi, tetas, tetav, wl, fi, c, wspd, azw, nr, ni, n12, w, wlp, ref_i, rwc, rw;
    float tds, tdv, rog, a, rwb;
    int iwl, iws1, iws2, isz1, isz2, ivz1, ivz2;
    pi = atanf(1.f) * 4.f;
    fac = pi / 180.f;
    paw = xpaw / fac;
    if (xphi < 0.0f) phi = -xphi;
    else if (xphi >= 2.0f * pi) phi = xphi - 2.0f * pi;
    else phi = xphi;
    if (xmuo <= 0.028f) tetas = acosf(0.028f) / fac; else tetas = acosf(xmuo) / fac;
    if (xmu <= 0.028f) tetav = acosf(0.028f) / fac; else tetav = acosf(xmu) / fac;
    if (pwl < 0.4f) wl = 0.4f; else if (pwl > 4.0f) wl = 4.0f; else wl = pwl;
    fi = 180.0f - phi / fac;
    c = pcl;
    wspd = fmaxf(0.25f, pws);
    azw = paw;
    indwat(wl, xsal, &nr, &ni);
    n12 = sqrtf(nr * nr + ni * ni);
    w = 2.95E-06f * powf(wspd, 3.52f);
    iwl = 1 + (int)((wl - 0.2f) / 0.1f);
    wlp = 0.5f + (iwl - 1) * 0.1f;
    ref_i = ref[iwl] + (wl - wlp) / 0.1f * (ref[iwl - 1] - ref[iwl]);
    rwc = w * ref_i;
    rw = morcasiwat(wl, c);
    getbound(wsbnd, 1, 6, wspd, &iws1, &iws2);
    getbound(angbnd, 1, 5, tetas, &isz1, &isz2);
    getbound(angbnd, 1, 5, tetav, &ivz1, &ivz2);
    tds = tdsbnd[isz1 - 1][iws1 - 1];
    tdv = tdvbnd[ivz1 - 1][iws1 - 1];
    rog = sunglint(wspd, nr, ni, azw, tetas, tetav, fi);
    a = 0.485f;
    rwb = (1 / (n12 * n12)) * tds * tdv * rw / (1 - a * rw);
    return rwc + (1 - w) * rog + (1 - rwc) * rwb;
}

/* SURFACE_BRDF  shdomsub2.f:1222-1301.  reflect is REFLECT(4,4) in Fortran order. */
int oracle_surface_brdf(int sfctype, const float *refparms, float wavelen, float mu2, float phi2,
                        float mu1, float phi1, int nstokes, float *reflect)
{
    const float pi = acosf(-1.0f);
    int i, j;
    if (sfctype == 'L' || sfctype == 'l') {
        for (j = 1; j <= nstokes; j++) for (i = 1; i <= nstokes; i++) RF(i, j) = 0.0f;
        RF(1, 1) = refparms[0];
    } else if (sfctype == 'W') {
        wave_fresnel_reflection(refparms[0], refparms[1], refparms[2], mu1, mu2, phi1, phi2, nstokes, reflect);
    } else if (sfctype == 'D') {
        diner_reflection(refparms[0], refparms[1], refparms[2], refparms[3], refparms[4],
                         -mu1, mu2, phi2 - phi1, nstokes, reflect);
    } else if (sfctype == 'R') {
        if (nstokes > 1) return 1;
        RF(1, 1) = rpv_reflection(refparms[0], refparms[1], refparms[2], -mu1, mu2, phi1 - phi2 - pi);
    } else if (sfctype == 'O') {
        if (nstokes > 1) return 1;
        RF(1, 1) = ocean_brdf_sw(refparms[0], -1.f, refparms[1], wavelen, -mu1, mu2, phi1 - phi2, phi1);
    } else if (sfctype == 'M') {
        float kgeo, kvol;
        if (nstokes > 1) return 1;
        ross_thick_li_sparse(refparms[0], refparms[1], refparms[2], 2.0f, 1.0f, -mu1, mu2, phi1 - phi2,
                             &RF(1, 1), &kgeo, &kvol);
    } else {
        return 1;
    }
    return 0;
}

/* VARIABLE_BRDF_SURFACE  shdomsub1.f:2597-2669.  bcrad_bot = BCRAD(:,1+NTOPPTS) viewed as
 * BCRAD(NSTOKES,NBOTPTS,*): slot 1 receives the upwelling radiance, slots 2.. hold the downwelling
 * radiances of the NANG/2 downward ordinates. */
int oracle_variable_brdf_surface(const oracle_state *st, int ibeg, int iend, float mu2, float phi2,
                                 float *bcrad_bot)
{
    const int ns = st->nstokes, nbot = st->nbotpts;
    const float opi = 1.0f / acosf(-1.0f);
    float reflect[16];
    int ibc, k, k1, jmu, jphi, jang;
#define BC(k, ibc, s) bcrad_bot[((k) - 1) + (size_t)ns * (((ibc) - 1) + (size_t)nbot * ((s) - 1))]
    memset(reflect, 0, sizeof(reflect));
    for (ibc = ibeg; ibc <= iend; ibc++) {
        const float *parms = st->sfcgridparms + (size_t)st->nsfcpar * (ibc - 1);
        for (k = 1; k <= ns; k++) BC(k, ibc, 1) = 0.0f;
        if (st->srctype != 'T') {
            const int i = st->bcptr[st->maxnbc + ibc - 1];
            if (oracle_surface_brdf(st->sfctype1, parms + 1, st->wavelen, mu2, phi2, st->solarmu, st->solaraz,
                                    ns, reflect)) return 1;
            for (k = 1; k <= ns; k++)
                BC(k, ibc, 1) = BC(k, ibc, 1) + opi * reflect[(k - 1)] * st->dirflux[i - 1];
        }
        jang = 1;
        for (jmu = 1; jmu <= st->nmu / 2; jmu++) {
            for (jphi = 1; jphi <= st->nphi0[jmu - 1]; jphi++) {
                float w;
                if (oracle_surface_brdf(st->sfctype1, parms + 1, st->wavelen, mu2, phi2, st->mu[jmu - 1],
                                        st->phi[(jmu - 1) + st->nmu * (jphi - 1)], ns, reflect)) return 1;
                w = opi * fabsf(st->mu[jmu - 1]) * st->wtdo[(jmu - 1) + st->nmu * (jphi - 1)];
                for (k1 = 1; k1 <= ns; k1++)
                    for (k = 1; k <= ns; k++)
                        BC(k, ibc, 1) = BC(k, ibc, 1) + w * reflect[(k - 1) + 4 * (k1 - 1)] * BC(k1, ibc, jang + 1);
                BC(1, ibc, 1) = BC(1, ibc, 1) + w * (1 - reflect[0]) * parms[0];
                for (k = 2; k <= ns; k++)
                    BC(k, ibc, 1) = BC(k, ibc, 1) - w * reflect[(k - 1)] * parms[0];
                jang = jang + 1;
            }
        }
    }
#undef BC
    return 0;
}
