// at3d_surface.cuh -- surface reflection models evaluated on the device (sm_100a).
// Replaces SURFACE_BRDF and the models it dispatches to (src/polarized/shdomsub2.f:1222-1699 of the AT3D
// reference: ROSS_THICK_LI_SPARSE, WAVE_FRESNEL_REFLECTION, DINER_REFLECTION, RPV_REFLECTION) and
// src/ocean_brdf.f (ocean_brdf_sw, morcasiwat, indwat, sunglint, Fresnel, getbound).  Declared precisions are kept (REAL -> float, REAL*8 -> double); the complex
// arithmetic of the Fresnel models is written out on (re, im) pairs.
#pragma once
#include "at3d_device.cuh"

struct cplx { double re, im; };
__device__ __forceinline__ cplx c_make(double r, double i) { cplx z; z.re = r; z.im = i; return z; }
__device__ __forceinline__ cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cplx c_scale(double s, cplx a) { return c_make(s * a.re, s * a.im); }
__device__ __forceinline__ cplx c_conj(cplx a) { return c_make(a.re, -a.im); }
__device__ __forceinline__ cplx c_div(cplx a, cplx b)
{
    const double d = b.re * b.re + b.im * b.im;
    return c_make((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
__device__ __forceinline__ double c_abs2(cplx a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ cplx c_sqrt(cplx a)
{   // principal branch
    const double m = sqrt(a.re * a.re + a.im * a.im);
    double re = sqrt(0.5 * (m + a.re)), im = sqrt(fmax(0.0, 0.5 * (m - a.re)));
    if (a.im < 0.0) im = -im;
    return c_make(re, im);
}

// REFLECT(4,4) in Fortran order; only the leading NSTOKES x NSTOKES block (NSTOKES <= 3) is set
#define RFL(i, j) reflect[((i) - 1) + 4 * ((j) - 1)]

// RPV_REFLECTION (shdomsub2.f:1661-1699)
__device__ __forceinline__ float dev_rpv_reflection(float rho0, float k, float theta, float mu1, float mu2, float phi)
{
    const float mu_min = 0.03f;
    float x1 = mu1, x2 = mu2;
    if (x1 < mu_min) x1 = mu_min;
    if (x2 < mu_min) x2 = mu_min;
    const float m = powf(x1 * x2 * (x1 + x2), k - 1);
    const float cosphi = cosf(phi);
    const float sin1 = sqrtf(1.0f - x1 * x1), sin2 = sqrtf(1.0f - x2 * x2);
    const float cosg = x1 * x2 + sin1 * sin2 * cosphi;
    const float f = (1 - theta * theta) / powf(1 + 2 * theta * cosg + theta * theta, 1.5f);
    const float tan1 = sin1 / x1, tan2 = sin2 / x2;
    const float capg = sqrtf(fabsf(tan1 * tan1 + tan2 * tan2 - 2 * tan1 * tan2 * cosphi));
    const float h = 1 + (1 - rho0) / (1 + capg);
    return rho0 * m * f * h;
}

// ROSS_THICK_LI_SPARSE (shdomsub2.f:1303-1357)
__device__ __forceinline__ float dev_ross_thick_li_sparse(float fiso, float fgeo, float fvol, float hb, float br,
                                                          float mudown, float muup, float relaz)
{
    const double pi = 3.14159265358979323846;
    double coseta = (double)(mudown * muup + sqrtf(1.0f - mudown * mudown) * sqrtf(1.0f - muup * muup) * cosf(relaz));
    const float kvol = (float)((((pi / 2.0 - acos(coseta)) * coseta + sqrt(1.0 - coseta * coseta))
                                / (double)(mudown + muup)) - pi / 4.0);
    const double theta_d = fabs(atan((double)br * sqrt(1.0 - (double)(mudown * mudown)) / (double)mudown));
    const double theta_u = fabs(atan((double)br * sqrt(1.0 - (double)(muup * muup)) / (double)muup));
    const double mu_d = cos(theta_d), mu_u = cos(theta_u);
    coseta = mu_d * mu_u + sqrt(1.0 - mu_d * mu_d) * sqrt(1.0 - mu_u * mu_u) * (double)cosf(relaz);
    const double sec_d = 1.0 / mu_d, sec_u = 1.0 / mu_u;
    const double tan_d = sqrt(1.0 - mu_d * mu_d) / mu_d, tan_u = sqrt(1.0 - mu_u * mu_u) / mu_u;
    const double dsq = tan_d * tan_d + tan_u * tan_u - 2 * tan_u * tan_d * (double)cosf(relaz);
    const double t = tan_d * tan_u * (double)sinf(relaz);
    double cost = (double)hb * sqrt(dsq + t * t) / (sec_d + sec_u);
    if (cost >= 1.0) cost = 1.0;
    if (cost <= -1.0) cost = -1.0;
    const double o = (acos(cost) - cost * sqrt(1.0 - cost * cost)) * (sec_d + sec_u) / pi;
    const float kgeo = (float)(o - sec_d - sec_u + 0.5 * (1.0 + coseta) * sec_u * sec_d);
    const float r = fiso + fvol * kvol + fgeo * kgeo;
    return fmaxf(r, 0.0f);
}

// WAVE_FRESNEL_REFLECTION (shdomsub2.f:1359-1520)
static __device__ void dev_wave_fresnel_reflection(float mre, float mim, float windspeed, float mui, float mur,
                                                   float phii, float phir, int nstokes, float *reflect)
{
    const cplx cn1 = c_make(1.0, 0.0), cn2 = c_make((double)mre, (double)mim);
    const double sigma2 = fmax(0.0005, 0.0015 + 0.00256 * (double)windspeed);
    double dmui = fabs((double)mui), dmur = (double)mur;
    if (fabs(dmui - 1.0) < 1e-10) dmui = 0.999999999999;
    if (fabs(dmur - 1.0) < 1e-10) dmur = 0.999999999999;
    const double dcosi = cos((double)phii), dsini = sin((double)phii);
    const double dcosr = cos((double)phir), dsinr = sin((double)phir);
    const double dsi = sqrt(1.0 - dmui * dmui), dsr = sqrt(1.0 - dmur * dmur);
    const double vi1 = dsi * dcosi, vi2 = dsi * dsini, vi3 = -dmui;
    const double vr1 = dsr * dcosr, vr2 = dsr * dsinr, vr3 = dmur;
    const double unit1 = vi1 - vr1, unit2 = vi2 - vr2, unit3 = vi3 - vr3;
    const double fact1 = unit1 * unit1 + unit2 * unit2 + unit3 * unit3;
    const double factor = sqrt(1.0 / fact1);
    const double xi1 = factor * (unit1 * vi1 + unit2 * vi2 + unit3 * vi3);
    const cplx cxi2 = c_sqrt(c_sub(c_make(1.0, 0.0),
                                   c_div(c_scale(1.0 - xi1 * xi1, c_mul(cn1, cn1)), c_mul(cn2, cn2))));
    cplx c1 = c_scale(xi1, cn1), c2 = c_mul(cn2, cxi2);
    const cplx crper = c_div(c_sub(c1, c2), c_add(c1, c2));
    c1 = c_scale(xi1, cn2); c2 = c_mul(cn1, cxi2);
    const cplx crpar = c_div(c_sub(c1, c2), c_add(c1, c2));
    const double ti1 = -dmui * dcosi, ti2 = -dmui * dsini, ti3 = -dsi;
    const double tr1 = dmur * dcosr, tr2 = dmur * dsinr, tr3 = -dsr;
    const double pi1 = -dsini, pi2 = dcosi, pr1 = -dsinr, pr2 = dcosr;
    const double pikr = pi1 * vr1 + pi2 * vr2;
    const double prki = pr1 * vi1 + pr2 * vi2;
    const double tikr = ti1 * vr1 + ti2 * vr2 + ti3 * vr3;
    const double trki = tr1 * vi1 + tr2 * vi2 + tr3 * vi3;
    const double e1 = pikr * prki, e2 = tikr * trki, e3 = tikr * prki, e4 = pikr * trki;
    const cplx cf11 = c_add(c_scale(e1, crper), c_scale(e2, crpar));
    const cplx cf12 = c_add(c_scale(-e3, crper), c_scale(e4, crpar));
    const cplx cf21 = c_add(c_scale(-e4, crper), c_scale(e3, crpar));
    const cplx cf22 = c_add(c_scale(e2, crper), c_scale(e1, crpar));
    const double vp1 = vi2 * vr3 - vi3 * vr2, vp2 = vi3 * vr1 - vi1 * vr3, vp3 = vi1 * vr2 - vi2 * vr1;
    double dmod = vp1 * vp1 + vp2 * vp2 + vp3 * vp3;
    dmod = dmod * dmod;
    const double rdz2 = unit3 * unit3, rdz4 = rdz2 * rdz2;
    const double dex = exp(-(unit1 * unit1 + unit2 * unit2) / (2 * sigma2 * rdz2));
    const double dcoeff = fact1 * fact1 * dex / (4 * dmui * dmur * dmod * rdz4 * 2 * sigma2);
    const double af = 0.5 * dcoeff;
    const double af11 = c_abs2(cf11), af12 = c_abs2(cf12), af21 = c_abs2(cf21), af22 = c_abs2(cf22);
    RFL(1, 1) = (float)((af11 + af12 + af21 + af22) * af);
    if (nstokes >= 2) {
        RFL(1, 2) = (float)((af11 - af12 + af21 - af22) * af);
        RFL(2, 1) = (float)((af11 - af22 + af12 - af21) * af);
        RFL(2, 2) = (float)((af11 - af12 - af21 + af22) * af);
    }
    if (nstokes >= 3) {
        const cplx c21 = c_conj(cf21), c22 = c_conj(cf22);
        const cplx ctttp = c_mul(cf11, c_conj(cf12)), cttpt = c_mul(cf11, c21), cttpp = c_mul(cf11, c22);
        const cplx ctppt = c_mul(cf12, c21), ctppp = c_mul(cf12, c22), cptpp = c_mul(cf21, c22);
        RFL(1, 3) = (float)((-ctttp.re - cptpp.re) * dcoeff);
        RFL(2, 3) = (float)((-ctttp.re + cptpp.re) * dcoeff);
        RFL(3, 1) = (float)((-cttpt.re - ctppp.re) * dcoeff);
        RFL(3, 2) = (float)((-cttpt.re + ctppp.re) * dcoeff);
        RFL(3, 3) = (float)((cttpp.re + ctppt.re) * dcoeff);
    }
    // shadowing
    const double p = 3.14159265358979323846;
    const double s1 = sqrt(2 * sigma2 / p), s3 = 1.0 / (sqrt(2 * sigma2)), s2 = s3 * s3;
    double dcot = dmui / sqrt(1.0 - dmui * dmui);
    double t1 = exp(-s2 * (dcot * dcot)), t2 = erfc(dcot * s3);
    const double shadowi = 0.5 * (s1 * t1 / dcot - t2);
    dcot = dmur / sqrt(1.0 - dmur * dmur);
    t1 = exp(-s2 * (dcot * dcot)); t2 = erfc(dcot * s3);
    const double shadowr = 0.5 * (s1 * t1 / dcot - t2);
    const double shadow = 1.0 / (1.0 + shadowi + shadowr);
    for (int j = 1; j <= nstokes; j++)
        for (int i = 1; i <= nstokes; i++) RFL(i, j) = (float)(RFL(i, j) * shadow);
}

// DINER_REFLECTION (shdomsub2.f:1524-1656).  The facet index of refraction is the real constant 1.5 there, so
// EPSILON, D, RP and RS are real (F34 = 0) and the COMPLEX expressions reduce to the REAL ones below.
static __device__ void dev_diner_reflection(float a, float k, float b, float zeta, float sigma,
                                            float mu1, float mu2, float phi, int nstokes, float *reflect)
{
    for (int j = 1; j <= nstokes; j++) for (int i = 1; i <= nstokes; i++) RFL(i, j) = 0.0f;
    const float sinth1 = sqrtf(1.0f - mu1 * mu1), sinth2 = sqrtf(1.0f - mu2 * mu2);
    const float cosphi = cosf(phi);
    float cosscatang = -mu1 * mu2 + sinth1 * sinth2 * cosphi;
    cosscatang = fminf(1.0f, fmaxf(-1.0f, cosscatang));
    const float tan1 = sinth1 / mu1, tan2 = sinth2 / mu2;
    const float capg = sqrtf(fabsf(tan1 * tan1 + tan2 * tan2 + 2 * tan1 * tan2 * cosphi));
    const float hot = 1 + (1 - a) / (1 + capg);
    RFL(1, 1) = a * powf((mu1 + mu2) * mu1 * mu2, k - 1) * expf(b * cosscatang);
    RFL(1, 1) = RFL(1, 1) * hot;
    if (zeta < 0.0f) return;
    const float gamma = 0.5f * acosf(-cosscatang);
    const float cosgamma = cosf(gamma);
    const float epsilon = 1.5f * 1.5f;
    const float d = sqrtf(epsilon - 1.0f + cosgamma * cosgamma);
    const float rp = (epsilon * cosgamma - d) / (epsilon * cosgamma + d);
    const float rs = (cosgamma - d) / (cosgamma + d);
    const float arp = fabsf(rp), ars = fabsf(rs);
    const float f11 = 0.5f * (arp * arp + ars * ars);
    const float f12 = 0.5f * (arp * arp - ars * ars);
    const float f33 = rp * rs;
    const float cosbeta = 0.5f * (mu1 + mu2) / cosgamma;
    float h;
    if (sigma > 0.0f) {
        const float cb2 = cosbeta * cosbeta;
        h = zeta * expf(-0.5f * (1 / cb2 - 1) / (sigma * sigma)) / (8 * (sigma * sigma) * mu2 * mu1 * (cb2 * cb2));
    } else {
        h = zeta / (8 * mu2 * mu1 * cosbeta);
    }
    RFL(1, 1) = RFL(1, 1) + h * f11;
    if (nstokes >= 2) {
        const float sinphi = sinf(phi);
        const float alpha1 = atanf(sinth2 * sinphi / (mu2 * sinth1 + sinth2 * mu1 * cosphi));
        const float alpha2 = atanf(sinth1 * sinphi / (sinth2 * mu1 + mu2 * sinth1 * cosphi));
        const float cos2alpha1 = cosf(2 * alpha1), sin2alpha1 = sinf(2 * alpha1);
        const float cos2alpha2 = cosf(2 * alpha2), sin2alpha2 = sinf(2 * alpha2);
        RFL(1, 2) = h * f12 * cos2alpha1;
        RFL(2, 1) = h * f12 * cos2alpha2;
        RFL(2, 2) = h * (f11 * cos2alpha1 * cos2alpha2 + f33 * sin2alpha1 * sin2alpha2);
        if (nstokes >= 3) {
            RFL(1, 3) = -h * f12 * sin2alpha1;
            RFL(2, 3) = h * (-f11 * sin2alpha1 * cos2alpha2 + f33 * cos2alpha1 * sin2alpha2);
            RFL(3, 1) = -h * f12 * sin2alpha2;
            RFL(3, 2) = h * (-f11 * cos2alpha1 * sin2alpha2 + f33 * sin2alpha1 * cos2alpha2);
            RFL(3, 3) = h * (f11 * sin2alpha1 * sin2alpha2 + f33 * cos2alpha1 * cos2alpha2);
        }
    }
}

// ---------------- ocean (src/ocean_brdf.f) ----------------
static __constant__ float c_oc_tkw[61] = {0.0209f,0.0200f,0.0196f,0.0189f,0.0183f,0.0182f,0.0171f,0.0170f,0.0168f,0.0166f,
    0.0168f,0.0170f,0.0173f,0.0174f,0.0175f,0.0184f,0.0194f,0.0203f,0.0217f,0.0240f,
    0.0271f,0.0320f,0.0384f,0.0445f,0.0490f,0.0505f,0.0518f,0.0543f,0.0568f,0.0615f,
    0.0640f,0.0640f,0.0717f,0.0762f,0.0807f,0.0940f,0.1070f,0.1280f,0.1570f,0.2000f,
    0.2530f,0.2790f,0.2960f,0.3030f,0.3100f,0.3150f,0.3200f,0.3250f,0.3300f,0.3400f,
    0.3500f,0.3700f,0.4050f,0.4180f,0.4300f,0.4400f,0.4500f,0.4700f,0.5000f,0.5500f,0.6500f};
static __constant__ float c_oc_txc[61] = {0.1100f,0.1110f,0.1125f,0.1135f,0.1126f,0.1104f,0.1078f,0.1065f,0.1041f,0.0996f,
    0.0971f,0.0939f,0.0896f,0.0859f,0.0823f,0.0788f,0.0746f,0.0726f,0.0690f,0.0660f,
    0.0636f,0.0600f,0.0578f,0.0540f,0.0498f,0.0475f,0.0467f,0.0450f,0.0440f,0.0426f,
    0.0410f,0.0400f,0.0390f,0.0375f,0.0360f,0.0340f,0.0330f,0.0328f,0.0325f,0.0330f,
    0.0340f,0.0350f,0.0360f,0.0375f,0.0385f,0.0400f,0.0420f,0.0430f,0.0440f,0.0445f,
    0.0450f,0.0460f,0.0475f,0.0490f,0.0515f,0.0520f,0.0505f,0.0440f,0.0390f,0.0340f,0.0300f};
static __constant__ float c_oc_te[61] = {0.668f,0.672f,0.680f,0.687f,0.693f,0.701f,0.707f,0.708f,0.707f,0.704f,
    0.701f,0.699f,0.700f,0.703f,0.703f,0.703f,0.703f,0.704f,0.702f,0.700f,
    0.700f,0.695f,0.690f,0.685f,0.680f,0.675f,0.670f,0.665f,0.660f,0.655f,
    0.650f,0.645f,0.640f,0.630f,0.623f,0.615f,0.610f,0.614f,0.618f,0.622f,
    0.626f,0.630f,0.634f,0.638f,0.642f,0.647f,0.653f,0.658f,0.663f,0.667f,
    0.672f,0.677f,0.682f,0.687f,0.695f,0.697f,0.693f,0.665f,0.640f,0.620f,0.600f};
static __constant__ float c_oc_tbw[61] = {0.0076f,0.0072f,0.0068f,0.0064f,0.0061f,0.0058f,0.0055f,0.0052f,0.0049f,0.0047f,
    0.0045f,0.0043f,0.0041f,0.0039f,0.0037f,0.0036f,0.0034f,0.0033f,0.0031f,0.0030f,
    0.0029f,0.0027f,0.0026f,0.0025f,0.0024f,0.0023f,0.0022f,0.0022f,0.0021f,0.0020f,
    0.0019f,0.0018f,0.0018f,0.0017f,0.0017f,0.0016f,0.0016f,0.0015f,0.0015f,0.0014f,
    0.0014f,0.0013f,0.0013f,0.0012f,0.0012f,0.0011f,0.0011f,0.0010f,0.0010f,0.0010f,
    0.0010f,0.0009f,0.0008f,0.0008f,0.0008f,0.0007f,0.0007f,0.0007f,0.0007f,0.0007f,0.0007f};
static __constant__ float c_oc_twl[62] = {0.250f,0.275f,0.300f,0.325f,0.345f,0.375f,0.400f,0.425f,0.445f,0.475f,
    0.500f,0.525f,0.550f,0.575f,0.600f,0.625f,0.650f,0.675f,0.700f,0.725f,
    0.750f,0.775f,0.800f,0.825f,0.850f,0.875f,0.900f,0.925f,0.950f,0.975f,
    1.000f,1.200f,1.400f,1.600f,1.800f,2.000f,2.200f,2.400f,2.600f,2.650f,
    2.700f,2.750f,2.800f,2.850f,2.900f,2.950f,3.000f,3.050f,3.100f,3.150f,
    3.200f,3.250f,3.300f,3.350f,3.400f,3.450f,3.500f,3.600f,3.700f,3.800f,3.900f,4.000f};
static __constant__ float c_oc_tnr[62] = {1.362f,1.354f,1.349f,1.346f,1.343f,1.341f,1.339f,1.338f,1.337f,1.336f,
    1.335f,1.334f,1.333f,1.333f,1.332f,1.332f,1.331f,1.331f,1.331f,1.330f,
    1.330f,1.330f,1.329f,1.329f,1.329f,1.328f,1.328f,1.328f,1.327f,1.327f,
    1.327f,1.324f,1.321f,1.317f,1.312f,1.306f,1.296f,1.279f,1.242f,1.219f,
    1.188f,1.157f,1.142f,1.149f,1.201f,1.292f,1.371f,1.426f,1.467f,1.483f,
    1.478f,1.467f,1.450f,1.432f,1.420f,1.410f,1.400f,1.385f,1.374f,1.364f,1.357f,1.351f};
static __constant__ float c_oc_tni[62] = {3.35E-08f,2.35E-08f,1.60E-08f,1.08E-08f,6.50E-09f,
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
static __constant__ float c_oc_ref[39] = {0.220f,0.220f,0.220f,0.220f,0.220f,0.220f,0.215f,0.210f,0.200f,0.190f,
    0.175f,0.155f,0.130f,0.080f,0.100f,0.105f,0.100f,0.080f,0.045f,0.055f,
    0.065f,0.060f,0.055f,0.040f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,
    0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f,0.000f};
static __constant__ float c_oc_angbnd[5] = {0.0f, 45.0f, 60.0f, 75.0f, 85.0f};
static __constant__ float c_oc_wsbnd[6] = {1.0f, 3.0f, 5.0f, 7.0f, 9.0f, 20.0f};
// DATA ((tdsbnd(iang,iws),iws=1,6),iang=1,5): rows are the angle bins
static __constant__ float c_oc_tds[5][6] = {
    {0.9787803f,0.9787738f,0.9787626f,0.9787467f,0.9787264f,0.9785573f},
    {0.9706900f,0.9698871f,0.9691746f,0.9685547f,0.9680276f,0.9666586f},
    {0.9479931f,0.9404608f,0.9385692f,0.9381815f,0.9384519f,0.9430056f},
    {0.9690591f,0.9275920f,0.9058769f,0.8951812f,0.8899654f,0.8892645f},
    {0.9980542f,0.9602273f,0.9114283f,0.8713799f,0.8417820f,0.7800314f}};
static __constant__ float c_oc_tdv[5][6] = {
    {0.9787764f,0.9787535f,0.9787106f,0.9786453f,0.9785548f,0.9775019f},
    {0.9692680f,0.9637051f,0.9564344f,0.9495727f,0.9438773f,0.9288712f},
    {0.9225163f,0.9069787f,0.9044844f,0.9052351f,0.9068328f,0.9153687f},
    {0.8048478f,0.8479503f,0.8678726f,0.8797889f,0.8878716f,0.9091171f},
    {0.7294627f,0.8137348f,0.8453338f,0.8629867f,0.8745421f,0.9036854f}};

// getbound (ocean_brdf.f:529-600) for the two ascending tables used; returns ind1 (1-based)
__device__ __forceinline__ int dev_getbound(const float *xvals, int ilast, float x)
{
    const int imid = ilast / 2 + 1;
    if (x > xvals[imid - 1]) {
        for (int i = imid; i <= ilast - 1; i++) if (xvals[i - 1] <= x && xvals[i] >= x) return i;
    } else {
        for (int i = 1; i <= imid; i++) if (xvals[i - 1] <= x && xvals[i] >= x) return i;
    }
    if (x < xvals[0]) return 1;
    return ilast - 1;
}

// morcasiwat (ocean_brdf.f:133-234)
static __device__ float dev_morcasiwat(float wl, float c)
{
    if (wl < 0.400f || wl > 0.700f) return 0.000f;
    const int iwl = 1 + (int)lroundf((wl - 0.400f) / 0.005f);
    const float kw = c_oc_tkw[iwl - 1], xc = c_oc_txc[iwl - 1], e = c_oc_te[iwl - 1], bw = c_oc_tbw[iwl - 1];
    float bb, kd;
    if (fabsf(c) < 0.0001f) { bb = 0.5f * bw; kd = kw; }
    else {
        const float b = 0.30f * powf(c, 0.62f);
        const float bbt = 0.002f + 0.02f * (0.5f - 0.25f * log10f(c)) * 0.550f / wl;
        bb = 0.5f * bw + bbt * b;
        kd = kw + xc * powf(c, e);
    }
    float r1 = 0.33f * bb / 0.75f / kd, r2;
    for (int it = 0; it < 200; it++) {
        const float u2 = 0.90f * (1.f - r1) / (1.f + 2.25f * r1);
        r2 = 0.33f * bb / u2 / kd;
        if (fabsf((r2 - r1) / r2) < 0.0001f) break;
        r1 = r2;
    }
    return r2;
}

// indwat (ocean_brdf.f:238-318)
__device__ __forceinline__ void dev_indwat(float wl, float xsal, float &nr, float &ni)
{
    int i = 2;
    while (!(wl < c_oc_twl[i - 1]) && i < 62) i++;
    const float xwl = c_oc_twl[i - 1] - c_oc_twl[i - 2];
    const float yr = c_oc_tnr[i - 1] - c_oc_tnr[i - 2], yi = c_oc_tni[i - 1] - c_oc_tni[i - 2];
    nr = c_oc_tnr[i - 2] + (wl - c_oc_twl[i - 2]) * yr / xwl;
    ni = c_oc_tni[i - 2] + (wl - c_oc_twl[i - 2]) * yi / xwl;
    const float nrc = 0.006f, nic = 0.000f;
    if (xsal >= 0.0f) { nr = nr + nrc * (xsal / 34.3f); ni = ni + nic * (xsal / 34.3f); }
    else { nr = nr + nrc; ni = ni + nic; }
}

// Fresnel (ocean_brdf.f:380-402)
__device__ __forceinline__ float dev_ocean_fresnel(float nr, float ni, float coschi, float sinchi)
{
    const float t = nr * nr - ni * ni - sinchi * sinchi;
    const float a1 = fabsf(t);
    const float a2 = sqrtf(t * t + 4 * nr * nr * ni * ni);
    const float u = sqrtf(0.5f * (a1 + a2));
    const float v = sqrtf(fmaxf(0.0f, 0.5f * (-a1 + a2)));
    const float rr2 = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    const float b1 = (nr * nr - ni * ni) * coschi, b2 = 2 * nr * ni * coschi;
    const float rl2 = ((b1 - u) * (b1 - u) + (b2 - v) * (b2 - v)) / ((b1 + u) * (b1 + u) + (b2 + v) * (b2 + v));
    return (rr2 + rl2) / 2.f;
}

// ocean_brdf_sw + sunglint (ocean_brdf.f:1-129, 322-377), split by what each sub-expression depends on so that
// the surface kernel evaluates every part once: OceanPoint = everything that depends on the surface parameters and the
// wavelength only (index of refraction, whitecaps, water-leaving reflectance, slope variances), OceanGeom = everything
// that depends on the incident/outgoing directions only (angles, facet slope, tilt, Fresnel angle).  dev_ocean_eval
// combines them; each arithmetic expression is the reference's.
struct OceanPoint { float nr, ni, n12, w, rwc, rw, sigmac, sigmau, c21, c03; int iws1; };
struct OceanGeom { float cs, cv, zx, zy, cphw, sphw, coschi, sinchi, ct4; int isz1, ivz1; };

static __device__ void dev_ocean_point(float pws, float xsal, float pcl, float pwl, OceanPoint &p)
{
    float wl;
    if (pwl < 0.4f) wl = 0.4f; else if (pwl > 4.0f) wl = 4.0f; else wl = pwl;
    const float wspd = fmaxf(0.25f, pws);
    dev_indwat(wl, xsal, p.nr, p.ni);
    p.n12 = sqrtf(p.nr * p.nr + p.ni * p.ni);
    p.w = 2.95E-06f * powf(wspd, 3.52f);
    const int iwl = 1 + (int)((wl - 0.2f) / 0.1f);
    const float wlp = 0.5f + (iwl - 1) * 0.1f;
    const float ref_i = c_oc_ref[iwl] + (wl - wlp) / 0.1f * (c_oc_ref[iwl - 1] - c_oc_ref[iwl]);
    p.rwc = p.w * ref_i;
    p.rw = dev_morcasiwat(wl, pcl);
    p.iws1 = dev_getbound(c_oc_wsbnd, 6, wspd);
    p.sigmac = 0.003f + 0.00192f * wspd;
    p.sigmau = 0.00316f * wspd;
    p.c21 = 0.01f - 0.0086f * wspd;
    p.c03 = 0.04f - 0.033f * wspd;
}

// The direction-only part splits once more: OceanInc depends on the incident direction alone (one per stored ordinate
// and state), OceanView on the outgoing direction alone (one per ray), and only the relative azimuth couples them.
struct OceanInc { float cs, ss, cphw, sphw; int isz1; };
struct OceanView { float cv, sv; int ivz1; };

__device__ __forceinline__ void dev_ocean_inc(float xmuo, float xpaw, OceanInc &q)
{
    const float pi = atanf(1.f) * 4.f, fac = pi / 180.f;
    const float paw = xpaw / fac;
    float tetas;
    if (xmuo <= 0.028f) tetas = acosf(0.028f) / fac; else tetas = acosf(xmuo) / fac;
    q.isz1 = dev_getbound(c_oc_angbnd, 5, tetas);
    const float phw = paw * fac;
    q.cs = cosf(tetas * fac); q.ss = sinf(tetas * fac);
    q.cphw = cosf(phw); q.sphw = sinf(phw);
}

__device__ __forceinline__ void dev_ocean_view(float xmu, OceanView &v)
{
    const float pi = atanf(1.f) * 4.f, fac = pi / 180.f;
    float tetav;
    if (xmu <= 0.028f) tetav = acosf(0.028f) / fac; else tetav = acosf(xmu) / fac;
    v.ivz1 = dev_getbound(c_oc_angbnd, 5, tetav);
    v.cv = cosf(tetav * fac); v.sv = sinf(tetav * fac);
}

__device__ __forceinline__ void dev_ocean_pair(const OceanInc &q, const OceanView &v, float xphi, OceanGeom &g)
{
    const float pi = atanf(1.f) * 4.f, fac = pi / 180.f;
    float phi;
    if (xphi < 0.0f) phi = -xphi;
    else if (xphi >= 2.0f * pi) phi = xphi - 2.0f * pi;
    else phi = xphi;
    const float fi = 180.0f - phi / fac;
    g.isz1 = q.isz1; g.ivz1 = v.ivz1;
    // sunglint(wspd, nr, ni, azw, ts, tv, fi)
    const float cs = q.cs, cv = v.cv, ss = q.ss, sv = v.sv;
    const float phir = fi * fac;
    g.cs = cs; g.cv = cv;
    g.zx = -sv * sinf(phir) / (cs + cv);
    g.zy = (ss + sv * cosf(phir)) / (cs + cv);
    const float tantilt = sqrtf(g.zx * g.zx + g.zy * g.zy);
    const float tilt = atanf(tantilt);
    g.cphw = q.cphw; g.sphw = q.sphw;
    float cos2chi = cv * cs + sv * ss * cosf(phir);
    if (cos2chi > 1.0f) cos2chi = 0.99999999999f;
    if (cos2chi < -1.0f) cos2chi = -0.99999999999f;
    g.coschi = sqrtf(0.5f * (1 + cos2chi));
    g.sinchi = sqrtf(0.5f * (1 - cos2chi));
    float ct = cosf(tilt);
    g.ct4 = (ct * ct) * (ct * ct);
}

static __device__ void dev_ocean_geom(float xmuo, float xmu, float xphi, float xpaw, OceanGeom &g)
{
    OceanInc q; OceanView v;
    dev_ocean_inc(xmuo, xpaw, q);
    dev_ocean_view(xmu, v);
    dev_ocean_pair(q, v, xphi, g);
}

__device__ __forceinline__ float dev_ocean_eval(const OceanPoint &p, const OceanGeom &g)
{
    const float pi = atanf(1.f) * 4.f;
    const float tds = c_oc_tds[g.isz1 - 1][p.iws1 - 1], tdv = c_oc_tdv[g.ivz1 - 1][p.iws1 - 1];
    const float c40 = 0.40f, c22 = 0.12f, c04 = 0.23f;
    const float xe = (g.cphw * g.zx + g.sphw * g.zy) / sqrtf(p.sigmac);
    const float xn = (-g.sphw * g.zx + g.cphw * g.zy) / sqrtf(p.sigmau);
    const float xe2 = xe * xe, xn2 = xn * xn;
    float coef = 1 - p.c21 / 2.f * (xe2 - 1) * xn - p.c03 / 6.f * (xn2 - 3) * xn;
    coef = coef + c40 / 24.f * (xe2 * xe2 - 6 * xe2 + 3);
    coef = coef + c04 / 24.f * (xn2 * xn2 - 6 * xn2 + 3);
    coef = coef + c22 / 4.f * (xe2 - 1) * (xn2 - 1);
    const float proba = coef / 2.f / pi / sqrtf(p.sigmau) / sqrtf(p.sigmac) * expf(-(xe2 + xn2) / 2.f);
    const float r1 = dev_ocean_fresnel(p.nr, p.ni, g.coschi, g.sinchi);
    const float rog = pi * r1 * proba / 4.f / g.cs / g.cv / g.ct4;
    const float a = 0.485f;
    const float rwb = (1 / (p.n12 * p.n12)) * tds * tdv * p.rw / (1 - a * p.rw);
    return p.rwc + (1 - p.w) * rog + (1 - p.rwc) * rwb;
}

static __device__ float dev_ocean_brdf_sw(float pws, float xsal, float pcl, float pwl, float xmuo, float xmu,
                                          float xphi, float xpaw)
{
    OceanPoint p; OceanGeom g;
    dev_ocean_point(pws, xsal, pcl, pwl, p);
    dev_ocean_geom(xmuo, xmu, xphi, xpaw, g);
    return dev_ocean_eval(p, g);
}

// SURFACE_BRDF (shdomsub2.f:1222-1301)
static __device__ void dev_surface_brdf(int sfctype, const float *refparms, float wavelen, float mu2, float phi2,
                                        float mu1, float phi1, int nstokes, float *reflect)
{
    if (sfctype == 'L' || sfctype == 'l') {
        for (int j = 1; j <= nstokes; j++) for (int i = 1; i <= nstokes; i++) RFL(i, j) = 0.0f;
        RFL(1, 1) = refparms[0];
    } else if (sfctype == 'W') {
        dev_wave_fresnel_reflection(refparms[0], refparms[1], refparms[2], mu1, mu2, phi1, phi2, nstokes, reflect);
    } else if (sfctype == 'D') {
        dev_diner_reflection(refparms[0], refparms[1], refparms[2], refparms[3], refparms[4],
                             -mu1, mu2, phi2 - phi1, nstokes, reflect);
    } else if (sfctype == 'R') {
        RFL(1, 1) = dev_rpv_reflection(refparms[0], refparms[1], refparms[2], -mu1, mu2, phi1 - phi2 - acosf(-1.0f));
    } else if (sfctype == 'O') {
        RFL(1, 1) = dev_ocean_brdf_sw(refparms[0], -1.f, refparms[1], wavelen, -mu1, mu2, phi1 - phi2, phi1);
    } else {   // 'M'
        RFL(1, 1) = dev_ross_thick_li_sparse(refparms[0], refparms[1], refparms[2], 2.0f, 1.0f, -mu1, mu2, phi1 - phi2);
    }
}
