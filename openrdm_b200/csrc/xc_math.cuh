// xc_math.cuh -- per-grid-point exchange-correlation algebra of OpenRDM's libfunc, as FP64
// device functions.  One function per reference entry point; each returns the *integrand*
// (energy per unit volume, i.e. what the reference multiplies by W(p) and accumulates):
//   ex_lsda_pt   Functional::EX_LSDA   /root/reference/src/libfunc/functional.cc:12-29
//   ex_pbe_pt    Functional::EX_PBE    functional.cc:31-99
//   ec_vwn3_pt   Functional::EC_VWN3   functional.cc:103-177
//   ec_pbe_pt    Functional::EC_PBE    functional.cc:179-309
// and translate_pt = MCPDFT::build_R + translate_density(+_gradients), mcpdft.cc:272-416.
//
// The branch structure (rho <= 1e-20, single-spin cases, sigma clamps, zeta/rs taken from the
// un-swapped density in EC_PBE) is the reference's.  The algebra is re-associated for the GPU:
// pow(x,4/3) -> x*cbrt(x), pow(x,2/3) -> cbrt(x)^2, shared sub-expressions are evaluated once
// (the reference evaluates A(r,zeta) three times and t() five times per point,
// functional.cc:268-271) and divisions are merged.  Every re-association is exact in real
// arithmetic; rounding differences are O(1e-16) relative per point.
#pragma once
#include <math.h>

namespace xc {

constexpr double TOL = 1.0e-20;  // functional.cc:39,106,185 ; mcpdft.cc:295
constexpr double PI = 3.14159265358979323846;
constexpr double C_LSDA = 0.9305257363491;       // 2^(1/3) * (9/8)(2/3)(3/pi)^(1/3), functional.cc:15-16,23
constexpr double THREE_PI2 = 29.608813203268074;  // 3 pi^2
constexpr double FZ_DEN = 0.5198420997897464;     // 2*2^(1/3) - 2, functional.cc:126
constexpr double PBE_BETA = 0.06672455060314922;  // functional.cc:37,208
constexpr double PBE_GAMMA = 0.0310906908696549;  // functional.cc:209
constexpr double PBE_KAPPA = 0.804;               // functional.cc:38
constexpr double PBE_MU = 0.219514972764517;      // beta pi^2 / 3, functional.cc:37
constexpr double D2FZ = 1.7099209341613656173;    // functional.cc:207

// ---------------------------------------------------------------- translation
struct Translated {
  double R, zeta;
  double ra, rb;           // translated spin densities, mcpdft.cc:326-330
  double saa, sab, sbb;    // translated sigma, mcpdft.cc:407-409 (zero when !GGA, energy.cc:63-65)
};

// build_R's formula, mcpdft.cc:282 (unguarded; inf/NaN are masked by the rho / Pi tests of the translation)
__device__ __forceinline__ double r_factor(double rho, double pi) { return 4.0 * pi / (rho * rho); }

// R is an argument because MCPDFT::translate reads the member R_ -- normally what build_R left there, but a caller
// may have installed its own through set_R (mcpdft.h:149)
template <bool GGA>
__device__ __forceinline__ Translated translate_pt(double rho, double pi, double gx, double gy, double gz, double R_in) {
  Translated t;
  t.R = R_in;
  const bool live = !(rho < TOL) && !(pi < 0.0);  // mcpdft.cc:319
  double z = 0.0;
  if (live && (1.0 - t.R) > TOL) z = sqrt(1.0 - t.R);  // mcpdft.cc:321-325
  t.zeta = z;
  const double up = live ? (1.0 + z) : 0.0, dn = live ? (1.0 - z) : 0.0;
  t.ra = up * (rho / 2.0);
  t.rb = dn * (rho / 2.0);
  if (GGA) {
    const double ax = up * (gx / 2.0), ay = up * (gy / 2.0), az = up * (gz / 2.0);  // mcpdft.cc:391-398
    const double bx = dn * (gx / 2.0), by = dn * (gy / 2.0), bz = dn * (gz / 2.0);
    t.saa = (ax * ax) + (ay * ay) + (az * az);
    t.sab = (ax * bx) + (ay * by) + (az * bz);
    t.sbb = (bx * bx) + (by * by) + (bz * bz);
  } else {
    t.saa = t.sab = t.sbb = 0.0;
  }
  return t;
}

// Fully-translated scheme, MCPDFT::fully_translate_density(+_gradients), mcpdft.cc:425-606:
// zeta(R) = sqrt(1-R) for R < R0, the quintic A dR^5 + B dR^4 + C dR^3 (dR = R - R1) on [R0, R1],
// 0 above R1; the translated gradients carry the derivative of zeta through grad R (px,py,pz =
// grad Pi).  pow(dR, k) of the reference becomes repeated multiplication.
constexpr double FT_R0 = 0.9, FT_R1 = 1.15;                                         // mcpdft.cc:427-428
constexpr double FT_A = -475.60656009, FT_B = -379.47331922, FT_C = -85.38149682;   // mcpdft.cc:429-431

template <bool GGA>
__device__ __forceinline__ Translated fully_translate_pt(double rho, double pi, double gx, double gy, double gz,
                                                         double px, double py, double pz, double R_in) {
  Translated t;
  t.R = R_in;
  const bool live = !(rho < TOL) && !(pi < 0.0);  // mcpdft.cc:458
  const double R = t.R, dR = R - FT_R1;
  const double d2 = dR * dR, d3 = d2 * dR, d4 = d2 * d2, d5 = d4 * dR;
  const bool lo = live && ((1.0 - R) > TOL) && (R < FT_R0);    // mcpdft.cc:460
  const bool mid = live && !lo && !(R < FT_R0) && !(R > FT_R1);  // mcpdft.cc:462
  const bool hi = live && !lo && !mid && (R > FT_R1);            // mcpdft.cc:464
  double z = 0.0;
  if (lo) z = sqrt(1.0 - R);
  else if (mid) z = FT_A * d5 + FT_B * d4 + FT_C * d3;
  t.zeta = z;
  t.ra = live ? (1.0 + z) * (rho / 2.0) : 0.0;  // mcpdft.cc:467-468
  t.rb = live ? (1.0 - z) * (rho / 2.0) : 0.0;
  t.saa = t.sab = t.sbb = 0.0;
  if (GGA) {
    const double g[3] = {gx, gy, gz}, q[3] = {px, py, pz};
    double a[3] = {0.0, 0.0, 0.0}, b[3] = {0.0, 0.0, 0.0};
    if (lo) {  // mcpdft.cc:541-548
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double u = (R * g[c]) / (2.0 * z) - q[c] / (rho * z);
        a[c] = (1.0 + z) * (g[c] / 2.0) + u;
        b[c] = (1.0 - z) * (g[c] / 2.0) - u;
      }
    } else if (mid) {  // mcpdft.cc:549-581
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double qr = q[c] / rho, rg = R * g[c];
        const double u = (FT_A * d4) * ((10.0 * qr) - (5.0 * rg)) + (FT_B * d3) * ((8.0 * qr) - (4.0 * rg)) +
                         (FT_C * d2) * ((6.0 * qr) - (3.0 * rg));
        a[c] = (1.0 + z) * (g[c] / 2.0) + u;
        b[c] = (1.0 - z) * (g[c] / 2.0) - u;
      }
    } else if (hi) {  // mcpdft.cc:582-590
#pragma unroll
      for (int c = 0; c < 3; ++c) { a[c] = g[c] / 2.0; b[c] = g[c] / 2.0; }
    }
    t.saa = (a[0] * a[0]) + (a[1] * a[1]) + (a[2] * a[2]);  // mcpdft.cc:599-601
    t.sab = (a[0] * b[0]) + (a[1] * b[1]) + (a[2] * b[2]);
    t.sbb = (b[0] * b[0]) + (b[1] * b[1]) + (b[2] * b[2]);
  }
  return t;
}

// ---------------------------------------------------------------- exchange
__device__ __forceinline__ double ex_lsda_pt(double ra, double rb) {
  // -(2^(1/3) Cx ra^(4/3) + 2^(1/3) Cx rb^(4/3)), no density threshold (functional.cc:22-27)
  return -(C_LSDA * (ra * cbrt(ra)) + C_LSDA * (rb * cbrt(rb)));
}

// rho_s * eX(2 rho_s) * FX(S(2 rho_s, 4 sigma_ss))   (functional.cc:41-60,89-90)
// kappa = 0.804 (PBE) or 1.245 (revPBE: the Psi4 plugin's EX_PBE_I, lsda_gga_functionals.cc:323)
__device__ __forceinline__ double pbe_x_spin(double rs_, double sig, double kappa) {
  const double r2 = 2.0 * rs_;
  const double kF = cbrt(THREE_PI2 * r2);
  const double eX = -(3.0 * kF) / (4.0 * PI);
  // s^2 = 4 sig / (2 kF r2)^2 = sig / (kF r2)^2 ;  FX = 1 + k - k^2 / (k + mu s^2)
  const double d = (kF * r2) * (kF * r2);
  const double FX = 1.0 + kappa - (kappa * kappa) * d / (kappa * d + PBE_MU * sig);
  return rs_ * eX * FX;
}

__device__ __forceinline__ double ex_pbe_pt(double ra, double rb, double saa, double sbb, double kappa = PBE_KAPPA) {
  const double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;                                     // functional.cc:74,94-96
  if (ra < TOL) return pbe_x_spin(rb, fmax(0.0, sbb), kappa);       // functional.cc:75-81
  if (rb < TOL) return pbe_x_spin(ra, fmax(0.0, saa), kappa);       // functional.cc:82-87
  return pbe_x_spin(ra, saa, kappa) + pbe_x_spin(rb, sbb, kappa);   // functional.cc:88-93
}

// ---- the Psi4 plugin's further functionals (src/libpsi4/lsda_gga_functionals.cc) ----
constexpr double REVPBE_KAPPA = 1.245;                 // lsda_gga_functionals.cc:323
constexpr double B88_BETA = 0.0042;                    // :212
constexpr double B88_C = 0.9305257363491001;           // 2^(1/3) * 0.73855876638202240586, :211,213

// -rho_s^(4/3) (c + beta X^2 / (1 + 6 beta X asinh X)),  X = sqrt(sigma_ss) / rho_s^(4/3)
__device__ __forceinline__ double b88_x_spin(double r, double sig) {
  const double r43 = r * cbrt(r);
  const double X = sqrt(sig) / r43;  // the reference reads the UNclamped sigma here (:237,245,253-254)
  return -r43 * (B88_C + (B88_BETA * (X * X)) / (1.0 + 6.0 * B88_BETA * X * asinh(X)));
}
// EX_B88_I, lsda_gga_functionals.cc:207-273
__device__ __forceinline__ double ex_b88_pt(double ra, double rb, double saa, double sbb) {
  const double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;
  if (ra < TOL) return b88_x_spin(rb, sbb);
  if (rb < TOL) return b88_x_spin(ra, saa);
  return b88_x_spin(ra, saa) + b88_x_spin(rb, sbb);
}

// EC_LYP_I, lsda_gga_functionals.cc:764-829: zero unless both spin densities exceed the threshold
__device__ __forceinline__ double ec_lyp_pt(double ra, double rb, double saa, double sab, double sbb) {
  const double rho = ra + rb;
  if (!(rho > TOL) || ra < TOL || rb < TOL) return 0.0;
  const double A = 0.04918, B = 0.132, C = 0.2533, Dd = 0.349;
  const double sigma = saa + sbb + 2.0 * sab;  // from the unclamped values (:787)
  const double sa = fmax(0.0, saa), sb = fmax(0.0, sbb);
  const double m13 = rcbrt(rho);               // rho^(-1/3)
  const double ca = cbrt(ra), cb = cbrt(rb);
  const double ca2 = ca * ca, cb2 = cb * cb, ca4 = ca2 * ca2, cb4 = cb2 * cb2;
  const double ra83 = ca4 * ca4, rb83 = cb4 * cb4;
  const double f = 1.0 / (1.0 + Dd * m13);
  const double m2 = m13 * m13, m4 = m2 * m2, m8 = m4 * m4;
  const double omega = exp(-C * m13) * f * (m8 * m2 * m13);  // rho^(-11/3)
  const double delta = (C + Dd * f) * m13;
  const double rho2 = rho * rho;
  return -4.0 * A * ra * rb * f / rho -
         A * B * omega *
             (ra * rb *
                  (36.462398978764777098 * (ra83 + rb83) + (2.6111111111111111111 - 0.38888888888888888889 * delta) * sigma -
                   (2.5 - 0.055555555555555555556 * delta) * (sa + sb) -
                   0.11111111111111111111 * (delta - 11.0) * ((ra * sa + rb * sb) / rho)) -
              0.66666666666666666667 * rho2 * sigma + (0.66666666666666666667 * rho2 - ra * ra) * sb +
              (0.66666666666666666667 * rho2 - rb * rb) * sa);
}

// EC_B88_OP, lsda_gga_functionals.cc:478-518.  No density threshold in the reference: 0/0 = NaN
// where a spin density vanishes, reproduced here.
__device__ __forceinline__ double b88op_K(double X) {
  return 1.8610514726982004 + 2.0 * (B88_BETA * (X * X)) / (1.0 + 6.0 * B88_BETA * X * asinh(X));  // 3 (3/4pi)^(1/3) + ...
}
__device__ __forceinline__ double ec_b88_op_pt(double ra, double rb, double saa, double sbb) {
  const double a13 = cbrt(ra), b13 = cbrt(rb);
  const double Ka = b88op_K(sqrt(saa) / (ra * a13)), Kb = b88op_K(sqrt(sbb) / (rb * b13));
  const double Bab = 2.3670 * (a13 * b13 * Ka * Kb) / (a13 * Ka + b13 * Kb);
  const double B2 = Bab * Bab;
  return -ra * rb * ((1.5214 * Bab + 0.5764) / (B2 * B2 + 1.1284 * (B2 * Bab) + 0.3183 * B2));
}

// ---------------------------------------------------------------- correlation
// f(zeta) = ((1+z)^(4/3) + (1-z)^(4/3) - 2) / (2 2^(1/3) - 2) from the two cube roots
__device__ __forceinline__ double spin_f(double z, double cu, double cd) {
  return ((1.0 + z) * cu + (1.0 - z) * cd - 2.0) / FZ_DEN;   // functional.cc:125-128
}

// VWN q(x) (functional.cc:137-142) with x = sqrt(rs), x^2 = rs; invQ = 1/sqrt(4d-c^2), Xp = X(p)
__device__ __forceinline__ double vwn_q(double x, double rs, double A, double p, double c, double d,
                                        double invQ, double Q, double Xp) {
  const double Xx = rs + c * x + d;
  const double at = atan(Q / (2.0 * x + c)) * invQ;
  const double xm = x - p;
  return A * (log(rs / Xx) + 2.0 * c * at - (c * p / Xp) * (log(xm * xm / Xx) + 2.0 * (c + 2.0 * p) * at));
}

__device__ __forceinline__ double ec_vwn3_pt(double ra, double rb) {
  double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;                    // functional.cc:159,171-174
  double zeta;
  if (ra < TOL) { rho = rb; zeta = 1.0; }          // functional.cc:160-162
  else if (rb < TOL) { rho = ra; zeta = 1.0; }     // functional.cc:163-165
  else zeta = (ra - rb) / rho;                     // functional.cc:167
  const double rs = cbrt(3.0 / (4.0 * PI * rho));  // functional.cc:121
  const double x = sqrt(rs);
  const double eP = vwn_q(x, rs, 0.03109070000, -0.409286, 13.0720, 42.7198, 22.271770159225095,
                          0.0448998886415768, 37.537128437796);
  const double eF = vwn_q(x, rs, 0.01554535000, -0.743294, 20.1231, 101.578, 0.8534715072594644,
                          1.1716852777089715, 87.17310647903601);
  const double fz = spin_f(zeta, cbrt(1.0 + zeta), cbrt(1.0 - zeta));
  return rho * (eP + fz * (eF - eP));              // functional.cc:169-170
}

// PW92 G (functional.cc:237-241) with sr = sqrt(r); p = 1 for all three parameter sets
__device__ __forceinline__ double pw92_G(double r, double sr, double T, double a1, double b1, double b2,
                                         double b3, double b4) {
  const double poly = T * (b1 * sr + b2 * r + b3 * (r * sr) + b4 * (r * r));
  return -2.0 * T * (1.0 + a1 * r) * log(1.0 + 0.5 / poly);
}

__device__ __forceinline__ double ec_pbe_pt(double ra, double rb, double saa, double sab, double sbb) {
  double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;                              // functional.cc:285,303-306
  double zeta = (ra - rb) / rho;                             // functional.cc:279 (before the branches)
  const double rs = cbrt(3.0 / (4.0 * PI * rho));            // functional.cc:280 (before the branches)
  double sig;
  if (ra < TOL) { rho = rb; sig = fmax(0.0, sbb); zeta = 1.0; }        // functional.cc:286-290
  else if (rb < TOL) { rho = ra; sig = fmax(0.0, saa); zeta = 1.0; }   // functional.cc:291-295
  else sig = fmax(0.0, saa) + fmax(0.0, sbb) + 2.0 * sab;              // functional.cc:296-300
  // PW92 eps_c(rs, zeta), functional.cc:243-261
  const double sr = sqrt(rs);
  const double mAc = pw92_G(rs, sr, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  const double eP = pw92_G(rs, sr, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const double eF = pw92_G(rs, sr, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const double cu = cbrt(1.0 + zeta), cd = cbrt(1.0 - zeta);
  const double fz = spin_f(zeta, cu, cd);
  const double z2 = zeta * zeta, z4 = z2 * z2;
  const double ec = eP + ((-mAc) * fz * (1.0 - z4)) / D2FZ + (eF - eP) * fz * z4;
  // H(rho, sigma, rs, zeta), functional.cc:210-233,263-271
  const double fi = 0.5 * (cu * cu + cd * cd);
  const double fi3 = fi * fi * fi;
  const double kF = cbrt(THREE_PI2 * rho);
  const double ks2 = 4.0 * kF / PI;
  const double t2 = sig / (4.0 * ks2 * (fi * fi) * (rho * rho));
  const double A = (PBE_BETA / PBE_GAMMA) / (exp(-ec / (fi3 * PBE_GAMMA)) - 1.0);
  const double At2 = A * t2;
  const double H = fi3 * PBE_GAMMA *
                   log(1.0 + (PBE_BETA / PBE_GAMMA) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  return rho * (H + ec);                                     // functional.cc:301-302
}

// EC_PW92_I, lsda_gga_functionals.cc:997-1093: the LSDA part of ec_pbe_pt (same branches)
__device__ __forceinline__ double ec_pw92_pt(double ra, double rb) {
  double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;
  double zeta = (ra - rb) / rho;
  const double rs = cbrt(3.0 / (4.0 * PI * rho));
  if (ra < TOL) { rho = rb; zeta = 1.0; }
  else if (rb < TOL) { rho = ra; zeta = 1.0; }
  const double sr = sqrt(rs);
  const double mAc = pw92_G(rs, sr, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  const double eP = pw92_G(rs, sr, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const double eF = pw92_G(rs, sr, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const double fz = spin_f(zeta, cbrt(1.0 + zeta), cbrt(1.0 - zeta));
  const double z2 = zeta * zeta, z4 = z2 * z2;
  return rho * (eP + ((-mAc) * fz * (1.0 - z4)) / D2FZ + (eF - eP) * fz * z4);
}

// EX_wPBE_I, rs_functionals.cc:86-300: short-range omega-PBE exchange (HSE exchange-hole enhancement
// factor), spin-scaled like EX_PBE_I.  R = 2 rho_s, SG = 4 sigma_ss, omega = MCPDFT_OMEGA.
__device__ __forceinline__ double wpbe_x_spin(double rs_, double sig, double omega) {
  const double A_bar = 0.757211, B = -0.106364, C = -0.118649, D = 0.609650, E = -0.0477963;
  const double R = 2.0 * rs_, SG = 4.0 * sig;
  const double kF = cbrt(THREE_PI2 * R);
  const double s1 = sqrt(SG) / (2.0 * kF * R);
  const double s2 = s1 * s1;
  // H(s) = (a2 s^2 + ... + a7 s^7) / (1 + b1 s + ... + b9 s^9), rs_functionals.cc:131-149 (Horner)
  const double Hn = s2 * (0.0159941 + s1 * (0.0852995 + s1 * (-0.160368 + s1 * (0.152645 + s1 * (-0.0971263 + s1 * 0.0422061)))));
  const double Hd = 1.0 + s1 * (5.33319 + s1 * (-12.4780 + s1 * (11.0988 + s1 * (-5.11013 + s1 * (1.71468 + s1 * (-0.610380 +
                    s1 * (0.307555 + s1 * (-0.0770547 + s1 * 0.0334840))))))));
  const double ze = s2 * (Hn / Hd);
  const double et = A_bar + ze, lam = D + ze;
  const double nu = omega / kF, nu2 = nu * nu;
  const double sq_la = sqrt(lam + nu2), sq_ze = sqrt(ze + nu2), sq_et = sqrt(et + nu2);
  const double chi = nu / sq_la;
  const double bg = lam / (1.0 - chi);
  const double Fb = 1.0 - (1.0 / (27.0 * C)) * (s2 / (1.0 + 0.25 * s2)) - (1.0 / (2.0 * C)) * ze;
  const double lam2 = lam * lam, lam3 = lam2 * lam, lam72 = lam3 * sqrt(lam);
  const double Gb = (-(2.0 / 5.0) * C * Fb * lam - (4.0 / 15.0) * B * lam2 - (6.0 / 5.0) * A_bar * lam3 -
                     (4.0 / 5.0) * 1.7724538509055160273 * lam72 - (12.0 / 5.0) * lam72 * (sqrt(ze) - sqrt(et))) * (1.0 / E);
  const double bg2 = bg * bg, bg3 = bg2 * bg;
  const double FX = A_bar * (ze / ((sq_ze + sq_et) * (sq_ze + nu)) + et / ((sq_ze + sq_et) * (sq_et + nu))) -
                    (4.0 / 9.0) * (B / bg) - (4.0 / 9.0) * (C * Fb / bg2) * (1.0 + 0.5 * chi) -
                    (8.0 / 9.0) * (E * Gb / bg3) * (1.0 + (9.0 / 8.0) * chi + (3.0 / 8.0) * (chi * chi)) +
                    2.0 * ze * log(1.0 - (lam - ze) / ((nu + sq_la) * (sq_la + sq_ze))) -
                    2.0 * et * log(1.0 - (lam - et) / ((nu + sq_la) * (sq_la + sq_et)));
  return rs_ * (-(3.0 * kF) / (4.0 * PI)) * FX;
}
__device__ __forceinline__ double ex_wpbe_pt(double ra, double rb, double saa, double sbb, double omega) {
  const double rho = ra + rb;
  if (!(rho > TOL)) return 0.0;
  if (ra < TOL) return wpbe_x_spin(rb, fmax(0.0, sbb), omega);
  if (rb < TOL) return wpbe_x_spin(ra, fmax(0.0, saa), omega);
  return wpbe_x_spin(ra, saa, omega) + wpbe_x_spin(rb, sbb, omega);
}

// ---------------------------------------------------------------- fused tPBE / trevPBE point
// EX_PBE + EC_PBE of one point with the cube roots shared between them (K3's hot case).  With both
// spin densities above the threshold,  2 rho_a = rho (1 + zeta)  and  2 rho_b = rho (1 - zeta),  so
//   kF(2 rho_s) = cbrt(3 pi^2 rho) * cbrt(1 +- zeta),    rs = cbrt(9 pi / 4) / cbrt(3 pi^2 rho):
// three cube roots instead of six and one division fewer; everything else is ex_pbe_pt / ec_pbe_pt
// term for term.  The single-spin and rho <= tol branches go through those two functions unchanged.
constexpr double C_RS_KF = 1.9191582926775128;  // cbrt(9 pi / 4) = rs * kF

__device__ __forceinline__ void pbe_pair_pt(double ra, double rb, double saa, double sab, double sbb, double kappa,
                                            double& ex, double& ec_out) {
  const double rho = ra + rb;
  if (!(rho > TOL) || ra < TOL || rb < TOL) {
    ex = ex_pbe_pt(ra, rb, saa, sbb, kappa);
    ec_out = ec_pbe_pt(ra, rb, saa, sab, sbb);
    return;
  }
  const double zeta = (ra - rb) / rho;
  const double k0 = cbrt(THREE_PI2 * rho);
  const double cu = cbrt(1.0 + zeta), cd = cbrt(1.0 - zeta);
  // ---- exchange (functional.cc:88-93): rho_s eX(2 rho_s) FX, FX = 1 + k - k^2 d / (k d + mu sigma), d = (kF 2 rho_s)^2
  {
    const double ka = k0 * cu, kb = k0 * cd;
    const double da = (ka * (2.0 * ra)) * (ka * (2.0 * ra)), db = (kb * (2.0 * rb)) * (kb * (2.0 * rb));
    const double na = kappa * da + PBE_MU * saa, nb = kappa * db + PBE_MU * sbb;
    const double k2 = kappa * kappa;
    // both quotients over one division
    const double inv = 1.0 / (na * nb);
    const double FXa = 1.0 + kappa - k2 * da * (nb * inv), FXb = 1.0 + kappa - k2 * db * (na * inv);
    const double c = -3.0 / (4.0 * PI);
    ex = ra * (c * ka) * FXa + rb * (c * kb) * FXb;
  }
  // ---- correlation (functional.cc:296-302; PW92 :243-261, H :210-233,263-271)
  // Seven of the reference's divisions share two reciprocals here (K3 is bound by its instruction count: ~12 instructions
  // per FP64 division): 1 / (k0 fi^3 rho^2) serves rs = C / k0, t^2 = sigma / (4 ks^2 fi^2 rho^2) and ec / (fi^3 gamma);
  // 1 / (p_A p_P p_F) serves the three 0.5 / poly of PW92's G; the constant divisors become multiplications.
  const double sig = fmax(0.0, saa) + fmax(0.0, sbb) + 2.0 * sab;
  const double fi = 0.5 * (cu * cu + cd * cd);
  const double fi2 = fi * fi, fi3 = fi2 * fi, rho2 = rho * rho;
  const double rD = 1.0 / (k0 * fi3 * rho2);
  const double rs = C_RS_KF * (rD * (fi3 * rho2));     // C / k0
  const double sr = sqrt(rs);
  const double rsr = rs * sr, rs2 = rs * rs;
  const double pA = 0.0168869 * (10.357 * sr + 3.6231 * rs + 0.88026 * rsr + 0.49671 * rs2);
  const double pP = 0.0310907 * (7.5957 * sr + 3.5876 * rs + 1.6382 * rsr + 0.49294 * rs2);
  const double pF = 0.01554535 * (14.1189 * sr + 6.1977 * rs + 3.3662 * rsr + 0.62517 * rs2);
  const double rp = 0.5 / (pA * pP * pF);
  const double mAc = -2.0 * 0.0168869 * (1.0 + 0.11125 * rs) * log(1.0 + rp * (pP * pF));
  const double eP = -2.0 * 0.0310907 * (1.0 + 0.21370 * rs) * log(1.0 + rp * (pA * pF));
  const double eF = -2.0 * 0.01554535 * (1.0 + 0.20548 * rs) * log(1.0 + rp * (pA * pP));
  const double fz = ((1.0 + zeta) * cu + (1.0 - zeta) * cd - 2.0) * (1.0 / FZ_DEN);
  const double z2 = zeta * zeta, z4 = z2 * z2;
  const double ec = eP + ((-mAc) * fz * (1.0 - z4)) * (1.0 / D2FZ) + (eF - eP) * fz * z4;
  const double t2 = sig * ((PI / 16.0) * (rD * fi));    // sigma / (4 ks^2 fi^2 rho^2), ks^2 = 4 k0 / pi
  const double A = (PBE_BETA / PBE_GAMMA) / (exp(-ec * (rD * (k0 * rho2)) * (1.0 / PBE_GAMMA)) - 1.0);
  const double At2 = A * t2;
  const double H = fi3 * PBE_GAMMA * log(1.0 + (PBE_BETA / PBE_GAMMA) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  ec_out = rho * (H + ec);
}

}  // namespace xc
