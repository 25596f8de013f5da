// -DREAL=double
// -DREAL_IS_F64=1
// -DNJ=6
// -DNC=53
// -DMODE=0
// -DNIN=3
// -DNOUT=6
// -DNRES=6
// -DPADIN=0
// -DMINB=5
// -DTPW=1

typedef REAL real;
typedef unsigned long long u64;
struct TrigC { real two_over_pi, magic, pio2_hi, pio2_mid, pio2_lo, fast_limit; real s[6]; real c[6]; };
struct SpecP { real C[NC]; real grav[3]; real fext[6]; real offset[NJ]; TrigC trig; };

// 1 where x > 0 (x < 0), else 0: compiles to a SET instruction, so the Coulomb term is two FMAs and no branch
__device__ __forceinline__ real step_pos(real x) { return x > (real)0 ? (real)1 : (real)0; }
__device__ __forceinline__ real step_neg(real x) { return x < (real)0 ? (real)1 : (real)0; }

// sincos of the NJ joint angles as one interleaved batch: three-FMA Cody-Waite reduction by pi/2, fdlibm minimax
// kernels on [-pi/4, pi/4], integer quadrant logic; every coefficient comes from the parameter bank (csrc/b2k_trig.cuh
// is the same code; measured <= 1.6 ulp).  fp32 rows with every |angle| < 8 take the special-function unit.
struct SC { real s, c; };
__device__ __noinline__ SC sincos_slow(real x)
{ // by value: taking the address of the caller's arrays would pin them to local memory on the fast path too
    SC r;
#if REAL_IS_F64
    sincos(x, &r.s, &r.c);
#else
    sincosf(x, &r.s, &r.c);
#endif
    return r;
}
__device__ __forceinline__ void sincos_batch(const real *x, const TrigC &t, real *s, real *c)
{
#if !REAL_IS_F64
    real amax = fabs(x[0]); // one comparison for the whole row (|x| is an operand modifier, max a single instruction)
#pragma unroll
    for (int j = 1; j < NJ; j++) amax = fmax(amax, fabs(x[j]));
    {
        if (amax < 8.0f) {
#pragma unroll
            for (int j = 0; j < NJ; j++) { s[j] = __sinf(x[j]); c[j] = __cosf(x[j]); }
            return;
        }
    }
#endif
#if REAL_IS_F64
    real amax = fabs(x[0]);
#pragma unroll
    for (int j = 1; j < NJ; j++) amax = fmax(amax, fabs(x[j]));
#endif
    if (!(amax < t.fast_limit)) { // rare: huge / non-finite angles somewhere in this row (NaN fails the test too)
#pragma unroll
        for (int j = 0; j < NJ; j++) { const SC r = sincos_slow(x[j]); s[j] = r.s; c[j] = r.c; }
        return;
    }
    real r[NJ], z[NJ], ps[NJ], pc[NJ];
    int q[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const real tt = fma(x[j], t.two_over_pi, t.magic);
#if REAL_IS_F64
        q[j] = __double2loint(tt);
#else
        q[j] = __float_as_int(tt);
#endif
        const real kd = tt - t.magic;
        real rr = fma(-kd, t.pio2_hi, x[j]);
        rr = fma(-kd, t.pio2_mid, rr);
        r[j] = fma(-kd, t.pio2_lo, rr);
        z[j] = r[j] * r[j];
    }
    const int D = REAL_IS_F64 ? 6 : 3;
#pragma unroll
    for (int j = 0; j < NJ; j++) { ps[j] = t.s[D - 1]; pc[j] = t.c[D - 1]; }
#pragma unroll
    for (int k = D - 2; k >= 0; k--) {
#pragma unroll
        for (int j = 0; j < NJ; j++) { ps[j] = fma(ps[j], z[j], t.s[k]); pc[j] = fma(pc[j], z[j], t.c[k]); }
    }
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const real sn = fma(r[j] * z[j], ps[j], r[j]);
        const real cs = fma(z[j] * z[j], pc[j], fma((real)-0.5, z[j], (real)1));
        const bool swap = q[j] & 1;
        const real ss = swap ? cs : sn;
        const real cc = swap ? sn : cs;
        const int sflip = (q[j] & 2) << 30;
        const int cflip = ((q[j] + 1) & 2) << 30;
#if REAL_IS_F64
        s[j] = __hiloint2double(__double2hiint(ss) ^ sflip, __double2loint(ss));
        c[j] = __hiloint2double(__double2hiint(cc) ^ cflip, __double2loint(cc));
#else
        s[j] = __int_as_float(__float_as_int(ss) ^ sflip);
        c[j] = __int_as_float(__float_as_int(cc) ^ cflip);
#endif
    }
}
__device__ __forceinline__ void rne_row(const real *C, const real *grav, const real *fext, const real *st, const real *ct, const real *qq, const real *qd, const real *qdd, real *out)
{
    const real t25 = qd[0] * C[0];
    const real t26 = qdd[0] * C[0];
    const real t27 = st[1] * qd[0];
    const real t28 = ct[1] * qd[0];
    const real t29 = qd[0] * qd[1];
    const real t30 = ct[1] * t29;
    const real t31 = fma(st[1], qdd[0], t30);
    const real t32 = -st[1] * t29;
    const real t33 = fma(ct[1], qdd[0], t32);
    const real t34 = st[1] * grav[2];
    const real t35 = ct[1] * grav[2];
    const real t36 = fma(qdd[1], C[1], t35);
    const real t37 = qd[1] * C[1];
    const real t38 = -t28 * C[1];
    const real t39 = fma(t28, t38, t34);
    const real t40 = fma(-qd[1], t37, t39);
    const real t41 = fma(-t27, t38, t36);
    const real t42 = t27 * t37;
    const real t43 = fma(-t33, C[1], t42);
    const real t44 = t28 * C[2];
    const real t45 = fma(-qd[1], C[3], t44);
    const real t46 = -qd[1] * C[4];
    const real t47 = fma(-t27, C[2], t46);
    const real t48 = t27 * C[3];
    const real t49 = fma(t28, C[4], t48);
    const real t50 = fma(t33, C[2], t40);
    const real t51 = fma(-qdd[1], C[3], t50);
    const real t52 = fma(-qdd[1], C[4], t41);
    const real t53 = fma(-t31, C[2], t52);
    const real t54 = fma(t31, C[3], t43);
    const real t55 = fma(t33, C[4], t54);
    const real t56 = fma(t28, t49, t51);
    const real t57 = fma(-qd[1], t47, t56);
    const real t58 = fma(qd[1], t45, t53);
    const real t59 = fma(-t27, t49, t58);
    const real t60 = fma(t27, t47, t55);
    const real t61 = fma(-t28, t45, t60);
    const real t62 = t57 * C[5];
    const real t63 = t59 * C[5];
    const real t64 = t61 * C[5];
    const real t65 = t27 * C[6];
    const real t66 = t28 * C[7];
    const real t67 = qd[1] * C[8];
    const real t68 = t28 * t67;
    const real t69 = fma(-qd[1], t66, t68);
    const real t70 = fma(t31, C[6], t69);
    const real t71 = qd[1] * t65;
    const real t72 = fma(-t27, t67, t71);
    const real t73 = fma(t33, C[7], t72);
    const real t74 = t27 * t66;
    const real t75 = fma(-t28, t65, t74);
    const real t76 = fma(qdd[1], C[8], t75);
    const real t77 = qd[1] + qd[2];
    const real t78 = ct[2] * t27;
    const real t79 = fma(st[2], t28, t78);
    const real t80 = -st[2] * t27;
    const real t81 = fma(ct[2], t28, t80);
    const real t82 = fma(t28, qd[2], t31);
    const real t83 = fma(-t27, qd[2], t33);
    const real t84 = qdd[1] + qdd[2];
    const real t85 = ct[2] * t82;
    const real t86 = fma(st[2], t83, t85);
    const real t87 = -st[2] * t82;
    const real t88 = fma(ct[2], t83, t87);
    const real t89 = ct[2] * t40;
    const real t90 = fma(st[2], t41, t89);
    const real t91 = -st[2] * t40;
    const real t92 = fma(ct[2], t41, t91);
    const real t93 = -t79 * C[9];
    const real t94 = fma(t77, C[10], t93);
    const real t95 = fma(t88, C[9], t90);
    const real t96 = fma(t88, C[10], -t43);
    const real t97 = fma(-t86, C[9], t92);
    const real t98 = fma(t84, C[10], t97);
    const real t99 = t81 * C[9];
    const real t100 = t81 * C[10];
    const real t101 = fma(-t77, t94, t95);
    const real t102 = fma(-t81, t100, t101);
    const real t103 = fma(t81, t99, t96);
    const real t104 = fma(-t79, t94, t103);
    const real t105 = fma(t79, t100, t98);
    const real t106 = fma(t77, t99, t105);
    const real t107 = -t77 * C[11];
    const real t108 = fma(t81, C[12], t107);
    const real t109 = -t81 * C[10];
    const real t110 = fma(-t79, C[11], t109);
    const real t111 = -t79 * C[12];
    const real t112 = fma(-t77, C[10], t111);
    const real t113 = fma(-t84, C[11], t102);
    const real t114 = fma(t88, C[12], t113);
    const real t115 = fma(-t88, C[10], t104);
    const real t116 = fma(-t86, C[11], t115);
    const real t117 = fma(-t86, C[12], t106);
    const real t118 = fma(-t84, C[10], t117);
    const real t119 = fma(-t77, t112, t114);
    const real t120 = fma(-t81, t110, t119);
    const real t121 = fma(t81, t108, t116);
    const real t122 = fma(-t79, t112, t121);
    const real t123 = fma(t79, t110, t118);
    const real t124 = fma(t77, t108, t123);
    const real t125 = t120 * C[13];
    const real t126 = t122 * C[13];
    const real t127 = t124 * C[13];
    const real t128 = t79 * C[14];
    const real t129 = -t77 * C[15];
    const real t130 = t81 * C[16];
    const real t131 = -t77 * t130;
    const real t132 = fma(-t81, t129, t131);
    const real t133 = fma(t86, C[14], t132);
    const real t134 = t81 * t128;
    const real t135 = fma(-t79, t130, t134);
    const real t136 = fma(-t84, C[15], t135);
    const real t137 = t79 * t129;
    const real t138 = fma(t77, t128, t137);
    const real t139 = fma(t88, C[16], t138);
    const real t140 = t81 + qd[3];
    const real t141 = ct[3] * t79;
    const real t142 = fma(-st[3], t77, t141);
    const real t143 = -st[3] * t79;
    const real t144 = fma(-ct[3], t77, t143);
    const real t145 = fma(-t77, qd[3], t86);
    const real t146 = fma(-t79, qd[3], -t84);
    const real t147 = t88 + qdd[3];
    const real t148 = ct[3] * t145;
    const real t149 = fma(st[3], t146, t148);
    const real t150 = -st[3] * t145;
    const real t151 = fma(ct[3], t146, t150);
    const real t152 = ct[3] * t102;
    const real t153 = fma(st[3], t104, t152);
    const real t154 = -st[3] * t102;
    const real t155 = fma(ct[3], t104, t154);
    const real t156 = fma(t151, C[1], t153);
    const real t157 = fma(t149, C[1], -t155);
    const real t158 = t144 * C[1];
    const real t159 = t142 * C[1];
    const real t160 = fma(t140, t159, t156);
    const real t161 = fma(-t144, t158, t106);
    const real t162 = fma(-t142, t159, t161);
    const real t163 = fma(-t140, t158, t157);
    const real t164 = fma(t151, C[17], t160);
    const real t165 = fma(t149, C[17], t163);
    const real t166 = t144 * C[17];
    const real t167 = t142 * C[17];
    const real t168 = fma(t140, t167, t164);
    const real t169 = fma(-t144, t166, t162);
    const real t170 = fma(-t142, t167, t169);
    const real t171 = fma(-t140, t166, t165);
    const real t172 = t168 * C[18];
    const real t173 = t170 * C[18];
    const real t174 = t171 * C[18];
    const real t175 = t142 * C[19];
    const real t176 = t140 * C[20];
    const real t177 = -t144 * C[19];
    const real t178 = t140 * t177;
    const real t179 = fma(t144, t176, t178);
    const real t180 = fma(t149, C[19], t179);
    const real t181 = -t144 * t175;
    const real t182 = fma(-t142, t177, t181);
    const real t183 = fma(t147, C[20], t182);
    const real t184 = t142 * t176;
    const real t185 = fma(-t140, t175, t184);
    const real t186 = fma(-t151, C[19], t185);
    const real t187 = -t144 + qd[4];
    const real t188 = ct[4] * t142;
    const real t189 = fma(st[4], t140, t188);
    const real t190 = -st[4] * t142;
    const real t191 = fma(ct[4], t140, t190);
    const real t192 = fma(t140, qd[4], t149);
    const real t193 = fma(-t142, qd[4], t147);
    const real t194 = -t151 + qdd[4];
    const real t195 = ct[4] * t192;
    const real t196 = fma(st[4], t193, t195);
    const real t197 = -st[4] * t192;
    const real t198 = fma(ct[4], t193, t197);
    const real t199 = ct[4] * t160;
    const real t200 = fma(st[4], t162, t199);
    const real t201 = -st[4] * t160;
    const real t202 = fma(ct[4], t162, t201);
    const real t203 = t200 * C[21];
    const real t204 = -t163 * C[21];
    const real t205 = t202 * C[21];
    const real t206 = t189 * C[22];
    const real t207 = -t187 * C[23];
    const real t208 = t191 * C[22];
    const real t209 = -t187 * t208;
    const real t210 = fma(-t191, t207, t209);
    const real t211 = fma(t196, C[22], t210);
    const real t212 = t191 * t206;
    const real t213 = fma(-t189, t208, t212);
    const real t214 = fma(-t194, C[23], t213);
    const real t215 = t189 * t207;
    const real t216 = fma(t187, t206, t215);
    const real t217 = fma(t198, C[22], t216);
    const real t218 = t191 + qd[5];
    const real t219 = ct[5] * t189;
    const real t220 = fma(-st[5], t187, t219);
    const real t221 = -st[5] * t189;
    const real t222 = fma(-ct[5], t187, t221);
    const real t223 = fma(-t187, qd[5], t196);
    const real t224 = fma(-t189, qd[5], -t194);
    const real t225 = t198 + qdd[5];
    const real t226 = ct[5] * t223;
    const real t227 = fma(st[5], t224, t226);
    const real t228 = -st[5] * t223;
    const real t229 = fma(ct[5], t224, t228);
    const real t230 = ct[5] * t200;
    const real t231 = fma(-st[5], t163, t230);
    const real t232 = -st[5] * t200;
    const real t233 = fma(-ct[5], t163, t232);
    const real t234 = fma(t229, C[24], t231);
    const real t235 = fma(-t227, C[24], t233);
    const real t236 = t222 * C[24];
    const real t237 = -t220 * C[24];
    const real t238 = fma(-t218, t237, t234);
    const real t239 = fma(t218, t236, t235);
    const real t240 = fma(t220, t237, t202);
    const real t241 = fma(-t222, t236, t240);
    const real t242 = t238 * C[25];
    const real t243 = t239 * C[25];
    const real t244 = t241 * C[25];
    const real t245 = t220 * C[26];
    const real t246 = t222 * C[26];
    const real t247 = t218 * C[27];
    const real t248 = t222 * t247;
    const real t249 = fma(-t218, t246, t248);
    const real t250 = fma(t227, C[26], t249);
    const real t251 = t218 * t245;
    const real t252 = fma(-t220, t247, t251);
    const real t253 = fma(t229, C[26], t252);
    const real t254 = t220 * t246;
    const real t255 = fma(-t222, t245, t254);
    const real t256 = fma(t225, C[27], t255);
    const real t257 = fma(-t243, C[24], t250);
    const real t258 = fma(t242, C[24], t253);
    const real t259 = fma(qdd[5], C[28], t256);
    const real t260 = fma(qd[5], C[29], t259);
    const real t261 = fma(C[31], step_pos(qd[5]), fma(-C[30], step_neg(qd[5]), t260));
    const real t262 = fma(ct[5], t242, t203);
    const real t263 = fma(-st[5], t243, t262);
    const real t264 = fma(st[5], t242, t204);
    const real t265 = fma(ct[5], t243, t264);
    const real t266 = t244 + t205;
    const real t267 = fma(ct[5], t257, t211);
    const real t268 = fma(-st[5], t258, t267);
    const real t269 = fma(st[5], t257, t214);
    const real t270 = fma(ct[5], t258, t269);
    const real t271 = t256 + t217;
    const real t272 = fma(qdd[4], C[32], -t270);
    const real t273 = fma(qd[4], C[33], t272);
    const real t274 = fma(C[35], step_pos(qd[4]), fma(-C[34], step_neg(qd[4]), t273));
    const real t275 = fma(t174, C[17], t180);
    const real t276 = fma(-t172, C[17], t186);
    const real t277 = fma(ct[4], t263, t172);
    const real t278 = fma(-st[4], t266, t277);
    const real t279 = fma(st[4], t263, t173);
    const real t280 = fma(ct[4], t266, t279);
    const real t281 = -t265 + t174;
    const real t282 = fma(t281, C[1], t275);
    const real t283 = fma(-t278, C[1], t276);
    const real t284 = fma(ct[4], t268, t282);
    const real t285 = fma(-st[4], t271, t284);
    const real t286 = fma(st[4], t268, t183);
    const real t287 = fma(ct[4], t271, t286);
    const real t288 = -t270 + t283;
    const real t289 = fma(qdd[3], C[36], t287);
    const real t290 = fma(qd[3], C[37], t289);
    const real t291 = fma(C[39], step_pos(qd[3]), fma(-C[38], step_neg(qd[3]), t290));
    const real t292 = fma(-t127, C[12], t133);
    const real t293 = fma(-t126, C[11], t292);
    const real t294 = fma(t125, C[11], t136);
    const real t295 = fma(t127, C[10], t294);
    const real t296 = fma(-t126, C[10], t139);
    const real t297 = fma(t125, C[12], t296);
    const real t298 = fma(ct[3], t278, t125);
    const real t299 = fma(st[3], t281, t298);
    const real t300 = fma(st[3], t278, t126);
    const real t301 = fma(-ct[3], t281, t300);
    const real t302 = t280 + t127;
    const real t303 = fma(-t302, C[9], t293);
    const real t304 = fma(-t302, C[10], t295);
    const real t305 = fma(t301, C[10], t297);
    const real t306 = fma(t299, C[9], t305);
    const real t307 = fma(ct[3], t285, t303);
    const real t308 = fma(st[3], t288, t307);
    const real t309 = fma(st[3], t285, t304);
    const real t310 = fma(-ct[3], t288, t309);
    const real t311 = t287 + t306;
    const real t312 = fma(qdd[2], C[40], -t310);
    const real t313 = fma(qd[2], C[41], t312);
    const real t314 = fma(C[43], step_pos(qd[2]), fma(-C[42], step_neg(qd[2]), t313));
    const real t315 = fma(t64, C[3], t70);
    const real t316 = fma(-t63, C[2], t315);
    const real t317 = fma(t62, C[2], t73);
    const real t318 = fma(t64, C[4], t317);
    const real t319 = fma(-t63, C[4], t76);
    const real t320 = fma(-t62, C[3], t319);
    const real t321 = fma(ct[2], t299, t62);
    const real t322 = fma(-st[2], t302, t321);
    const real t323 = fma(st[2], t299, t63);
    const real t324 = fma(ct[2], t302, t323);
    const real t325 = -t301 + t64;
    const real t326 = fma(-t325, C[1], t318);
    const real t327 = fma(t324, C[1], t320);
    const real t328 = fma(ct[2], t308, t316);
    const real t329 = fma(-st[2], t311, t328);
    const real t330 = fma(st[2], t308, t326);
    const real t331 = fma(ct[2], t311, t330);
    const real t332 = -t310 + t327;
    const real t333 = fma(qdd[1], C[44], t332);
    const real t334 = fma(qd[1], C[45], t333);
    const real t335 = fma(C[47], step_pos(qd[1]), fma(-C[46], step_neg(qd[1]), t334));
    const real t336 = ct[1] * t322;
    const real t337 = fma(-st[1], t324, t336);
    const real t338 = st[1] * t322;
    const real t339 = fma(ct[1], t324, t338);
    const real t340 = ct[1] * t329;
    const real t341 = fma(-st[1], t331, t340);
    const real t342 = fma(t325, C[48], t341);
    const real t343 = fma(st[1], t329, t26);
    const real t344 = fma(ct[1], t331, t343);
    const real t345 = fma(-t337, C[48], t332);
    const real t346 = fma(qdd[0], C[49], t344);
    const real t347 = fma(qd[0], C[50], t346);
    const real t348 = fma(C[52], step_pos(qd[0]), fma(-C[51], step_neg(qd[0]), t347));
    out[0] = t348;
    out[1] = t335;
    out[2] = t314;
    out[3] = t291;
    out[4] = t274;
    out[5] = t261;
}

#define LDI (PADIN ? (NJ | 1) : NJ)                  /* smem row stride of the input tiles, in reals */
#define IN_BYTES ((32 * LDI * (int)sizeof(real) + 15) & ~15)
#define OUT_BYTES (32 * NOUT * (int)sizeof(real))
#define NBUF (TPW > 1 ? 2 : 1)                        /* input buffers per warp */
#define WARP_BYTES (NBUF * NIN * IN_BYTES + OUT_BYTES)

__device__ __forceinline__ void load_tile(real *s, const real *g, int lane)
{
#if PADIN
    // padded rows: the exact image would make the one-row-per-lane reads collide on the shared-memory banks
    for (int i = lane; i < 32 * NJ; i += 32) {
        const int r = i / NJ, c = i - r * NJ;
        s[r * LDI + c] = g[i];
    }
#else
    const uint4 *gg = reinterpret_cast<const uint4 *>(g) + lane;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s) + 16u * (unsigned)lane;
    constexpr int UNITS = (32 * NJ * (int)sizeof(real)) / 16; // 16-byte units in the tile: trip count known at compile time
#pragma unroll
    for (int k = 0; k < (UNITS + 31) / 32; k++)
        if (k * 32 + 32 <= UNITS || lane < UNITS - k * 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa + 512u * (unsigned)k), "l"(gg + 32 * k));
#endif
}

// A warp owns TPW consecutive tiles of 32 rows (one-shot grid of full tiles; the ragged tail of a batch goes to the
// generic kernel).  The input tiles are double-buffered: the cp.async loads of tile t+1 are issued before tile t is
// computed, so a warp always has a tile's worth of reads in flight -- with one tile per warp and 16-20 resident warps
// per SM the kernel was latency-bound (ncu: long-scoreboard the top stall, FP64 pipe 62 % busy; profiles/r02_rne64s_v1.txt).
extern "C" __global__ void __launch_bounds__(128, MINB)
k_rne_spec(const __grid_constant__ SpecP P, const real *__restrict__ in0, const real *__restrict__ in1,
           const real *__restrict__ in2, real *__restrict__ out, long long ntiles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long tile0 = ((long long)blockIdx.x * 4 + warp) * TPW;
    if (tile0 >= ntiles) return;
    unsigned char *wb = smem + (size_t)warp * WARP_BYTES;
    real *so = reinterpret_cast<real *>(wb + NBUF * NIN * IN_BYTES);
    auto load = [&](long long tile, int buf) {
        unsigned char *b = wb + (size_t)buf * NIN * IN_BYTES;
        const size_t row0 = (size_t)tile * 32;
        load_tile(reinterpret_cast<real *>(b), in0 + row0 * NJ, lane);
        if (NIN >= 2) load_tile(reinterpret_cast<real *>(b + IN_BYTES), in1 + row0 * NJ, lane);
        if (NIN >= 3) load_tile(reinterpret_cast<real *>(b + 2 * IN_BYTES), in2 + row0 * NJ, lane);
#if !PADIN
        asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
    };
    load(tile0, 0);
#pragma unroll 1
    for (int t = 0; t < TPW; t++) {
        const long long tile = tile0 + t;
        if (tile >= ntiles) break;
        const bool more = (t + 1 < TPW) && (tile + 1 < ntiles);
        if (more) load(tile + 1, (t + 1) & (NBUF - 1));
#if !PADIN
        if (more) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
        __syncwarp();
        const unsigned char *b = wb + (size_t)(t & (NBUF - 1)) * NIN * IN_BYTES;
        const real *s0 = reinterpret_cast<const real *>(b);
        const real *s1 = reinterpret_cast<const real *>(b + IN_BYTES);
        const real *s2 = reinterpret_cast<const real *>(b + 2 * IN_BYTES);
        real th[NJ], st[NJ], ct[NJ], a1[NJ], a2[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            th[j] = s0[lane * LDI + j] + P.offset[j];
            a1[j] = NIN >= 2 ? s1[lane * LDI + j] : (real)0;
            a2[j] = NIN >= 3 ? s2[lane * LDI + j] : (real)0;
        }
        sincos_batch(th, P.trig, st, ct);
        real res[NRES];
        rne_row(P.C, P.grav, P.fext, st, ct, th, a1, a2, res);
#if MODE == 5
        // accel: res = [M (NJ x NJ, row i = torques for a unit acceleration of joint i) | torque - rne(q, qd, 0)].
        // M is the joint-space inertia matrix (symmetric positive definite): LDL^T without pivoting, in registers.
        real d[NJ];
#pragma unroll
        for (int c = 0; c < NJ; c++) {
            real dc = res[c * NJ + c];
#pragma unroll
            for (int k = 0; k < c; k++) dc = fma(-res[c * NJ + k] * d[k], res[c * NJ + k], dc);
            d[c] = dc;
            const real inv = (real)1 / dc;
#pragma unroll
            for (int r = c + 1; r < NJ; r++) {
                real v = res[r * NJ + c];
#pragma unroll
                for (int k = 0; k < c; k++) v = fma(-res[r * NJ + k] * d[k], res[c * NJ + k], v);
                res[r * NJ + c] = v * inv; // L[r][c]
            }
        }
        real *y = res + NJ * NJ;
#pragma unroll
        for (int r = 0; r < NJ; r++)
#pragma unroll
            for (int k = 0; k < r; k++) y[r] = fma(-res[r * NJ + k], y[k], y[r]);
#pragma unroll
        for (int r = 0; r < NJ; r++) y[r] = y[r] / d[r];
#pragma unroll
        for (int r = NJ - 1; r >= 0; r--)
#pragma unroll
            for (int k = r + 1; k < NJ; k++) y[r] = fma(-res[k * NJ + r], y[k], y[r]);
        const real *o = y;
#else
        const real *o = res;
#endif
        // the previous tile's bulk copy must have finished READING the stage before it is overwritten
        if (t > 0) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < NOUT; k++) so[lane * NOUT + k] = o[k];
        // the staged tile is the exact image of the output block: one TMA bulk copy (shared -> global) by lane 0
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const unsigned ss = (unsigned)__cvta_generic_to_shared(so);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)tile * 32 * NOUT), "r"(ss),
                         "r"((unsigned)OUT_BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the copies read this warp's shared memory
}
