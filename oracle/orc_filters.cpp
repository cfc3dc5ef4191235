// orc_filters.cpp — see orc_filters.h.  TEST INFRASTRUCTURE ONLY.
#include "orc_filters.h"

namespace orc {

namespace {

struct Img16 {
    const uint16_t* p;
    int W, H;
    f4 load(int x, int y) const
    {
        if (x < 0 || y < 0 || x >= W || y >= H || !p) return f4(0, 0, 0, 0);
        const uint16_t* q = p + 4 * ((size_t)y * W + x);
        return f4(half_to_float(q[0]), half_to_float(q[1]), half_to_float(q[2]), half_to_float(q[3]));
    }
    f3 rgb(int x, int y) const { return load(x, y).xyz(); }
};
struct ImgU {
    const uint32_t* p;
    int W, H;
    uint32_t load(int x, int y) const
    {
        if (x < 0 || y < 0 || x >= W || y >= H || !p) return 0;
        return p[(size_t)y * W + x];
    }
};

inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline f3 mix3(f3 a, f3 b, float t) { return f3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
inline f3 rgb2ycocg(f3 c) { return f3(0.25f * c.x + 0.5f * c.y + 0.25f * c.z, 0.5f * c.x - 0.5f * c.z, -0.25f * c.x + 0.5f * c.y - 0.25f * c.z); }
inline f3 ycocg2rgb(f3 c) { return f3(c.x + c.y - c.z, c.x + c.z, c.x - c.y - c.z); }
inline f3 clamp3(f3 v, f3 lo, f3 hi) { return vmin(vmax(v, lo), hi); }
inline void store(uint16_t* out, int W, int H, int x, int y, f3 rgb, float a)
{
    if (x < 0 || y < 0 || x >= W || y >= H) return;
    uint16_t* q = out + 4 * ((size_t)y * W + x);
    q[0] = float_to_half_rne(rgb.x), q[1] = float_to_half_rne(rgb.y), q[2] = float_to_half_rne(rgb.z), q[3] = float_to_half_rne(a);
}

float calculateWeight(float centerDist, bool sameObject, bool isCenter, f3 n, f3 cn) // ReProject:40-55
{
    if (isCenter) return 0.4f;
    if (!sameObject) return 0.0f;
    float nd = clampf(dot(n, cn), 0.0f, 1.0f);
    const float th = 0.98f;
    if (nd < th) return 0.0f;
    const float nw = (nd - th) / (1.0f - th);
    return nw * 2.0f / (centerDist * 1.5f + 4.0f);
}

// Const_Func.slang:84-121
float W_f(float x, float e0, float e1)
{
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    float a = (x - e0) / (e1 - e0);
    return a * a * (3 - 2 * a);
}
float H_f(float x, float e0, float e1)
{
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    return (x - e0) / (e1 - e0);
}
float granTurismo(float x)
{
    const float e = 2.71828f;
    float P = 1, a = 0.7f, m = 0.22f, l = 0.4f, c = 1.33f, b = 0;
    float l0 = (P - m) * l / a;
    float L_x = m + a * (x - m);
    float T_x = m * powf(x / m, c) + b;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = a * P / (P - S1);
    float S_x = P - (P - S1) * powf(e, -(C2 * (x - S0) / P));
    float w0 = 1 - W_f(x, 0, m);
    float w2 = H_f(x, m + l0, m + l0);
    float w1 = 1 - w0 - w2;
    return T_x * w0 + L_x * w1 + S_x * w2;
}
f3 gt3(f3 v) { return f3(granTurismo(v.x), granTurismo(v.y), granTurismo(v.z)); }
f3 st2084(f3 lin) // Const_Func.slang:51-68
{
    const float m1 = 0.1593017578125f, m2 = 78.84375f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f, C = 10000.f;
    auto one = [&](float v) {
        float L = v / C;
        float Lm = powf(L, m1);
        float N1 = c1 + c2 * Lm, N2 = 1.0f + c3 * Lm;
        float N = N1 * (1.0f / N2);
        return powf(N, m2);
    };
    return f3(one(lin.x), one(lin.y), one(lin.z));
}

bool edgeDetect(uint32_t center, const ImgU& im, int x, int y) // DenoiseJBF:55-68
{
    uint32_t a = im.load(x + 1, y + 1), b = im.load(x - 1, y - 1), c = im.load(x - 1, y + 1), d = im.load(x + 1, y - 1);
    bool e0 = a != center || b != center || c != center || d != center;
    bool e1 = a == center || b == center || c == center || d == center;
    return e0 && e1;
}

} // namespace

void reproject(const GkUniformBufferObject& U, uint32_t W_, uint32_t H_, bool needClamp, bool /*needSpatio: unused by the shader*/,
               const uint16_t* src_, const uint16_t* hist_, const float* motion_, const uint32_t* id0_, const uint32_t* id1_,
               const uint16_t* normal_, uint16_t* out)
{
    const int W = (int)W_, H = (int)H_;
    const Img16 Src{src_, W, H}, Hist{hist_, W, H}, Nrm{normal_, W, H};
    const ImgU Id0{id0_, W, H}, Id1{id1_, W, H};
    const int vx = (int)U.ViewportRect[0], vy = (int)U.ViewportRect[1];
    const int vEndX = (int)(U.ViewportRect[0] + U.ViewportRect[2]), vEndY = (int)(U.ViewportRect[1] + U.ViewportRect[3]);
    for (int ty = 0; ty < H; ++ty)
        for (int tx = 0; tx < W; ++tx) {
            const int x = tx + vx, y = ty + vy;
            if (x >= W || y >= H) continue;
            const f4 src = Src.load(x, y);
            const f2 motion = (x < W && y < H) ? f2(motion_[2 * ((size_t)y * W + x)], motion_[2 * ((size_t)y * W + x) + 1]) : f2(0, 0);
            const int px = (int)floorf(float(x) + motion.x), py = (int)floorf(float(y) + motion.y);
            const bool inside = (px < vEndX && py < vEndY) && (px >= vx - 1 && py >= vy - 1);
            const uint32_t cur0 = Id0.load(x, y);
            uint32_t p0 = Id1.load(px, py);
            f3 fin = src.xyz();
            if (U.ProgressiveRender) {
                const f4 h = Hist.load(x, y);
                const float t = clampf(1.0f / float(U.TemporalFrames), 0.0f, 1.0f);
                store(out, W, H, x, y, mix3(h.xyz(), src.xyz(), t), 1.0f);
                continue;
            }
            bool useHistory = true;
            if (cur0 == 65535 || U.TotalFrames == 0 || !inside) useHistory = false;
            if (useHistory) {
                const int R = U.DisableSpatialReuse ? 0 : 2;
                const f3 cn = Nrm.rgb(x, y);
                f4 spatial(0, 0, 0, 0);
                float total = 0;
                for (int dy = -R; dy <= R; ++dy)
                    for (int dx = -R; dx <= R; ++dx) {
                        const f4 sp = Src.load(x + dx, y + dy);
                        const uint32_t pid = Id0.load(x + dx, y + dy);
                        const f3 wn = Nrm.rgb(x + dx, y + dy);
                        const float cd = sqrtf(float(dx) * float(dx) + float(dy) * float(dy));
                        const float w = calculateWeight(cd, pid == cur0, dx == 0 && dy == 0, wn, cn);
                        spatial = spatial + sp * w;
                        total += w;
                    }
                spatial = spatial / total;
                uint32_t p1 = Id1.load(px + 1, py), p2 = Id1.load(px, py + 1), p3 = Id1.load(px + 1, py + 1);
                if (length(motion) < 0.02f) p0 = p1 = p2 = p3 = cur0;
                f3 hc[4];
                hc[0] = cur0 == p0 ? Hist.rgb(px, py) : spatial.xyz();
                hc[1] = cur0 == p1 ? Hist.rgb(px + 1, py) : spatial.xyz();
                hc[2] = cur0 == p2 ? Hist.rgb(px, py + 1) : spatial.xyz();
                hc[3] = cur0 == p3 ? Hist.rgb(px + 1, py + 1) : spatial.xyz();
                const float fx = (float(x) + motion.x) - floorf(float(x) + motion.x), fy = (float(y) + motion.y) - floorf(float(y) + motion.y);
                f3 history = mix3(mix3(hc[0], hc[1], fx), mix3(hc[2], hc[3], fx), fy);
                history = clamp3(history, f3(0.0f), f3(1600.0f));
                if (needClamp) {
                    f3 mx = rgb2ycocg(src.xyz()), mn = mx;
                    for (int k = 0; k < 25; ++k) {
                        const f3 c = rgb2ycocg(Src.rgb(x + (k % 5 - 2), y + (k / 5 - 2)));
                        mn = vmin(mn, c), mx = vmax(mx, c);
                    }
                    history = ycocg2rgb(clamp3(rgb2ycocg(history), mn, mx));
                }
                const uint32_t tf = U.TemporalFrames > 1 ? U.TemporalFrames : 1;
                const float keep = 1.0f / float(tf);
                fin = mix3(history, src.xyz(), clampf(keep, 0.0f, 1.0f));
            }
            store(out, W, H, x, y, fin, 1.0f);
        }
}

void denoiseJBF(const GkUniformBufferObject& U, uint32_t W_, uint32_t H_, const uint16_t* diffuse_, const uint16_t* spec_,
                const uint16_t* normal_, const uint32_t* id0_, const uint32_t* id1_, const uint16_t* albedo_, uint16_t* out)
{
    const int W = (int)W_, H = (int)H_;
    const Img16 Dif{diffuse_, W, H}, Spc{spec_, W, H}, Alb{albedo_, W, H};
    (void)normal_;
    const ImgU Id0{id0_, W, H}, Id1{id1_, W, H};
    const int vx = (int)U.ViewportRect[0], vy = (int)U.ViewportRect[1];
    const f3 lumW(0.212671f, 0.715160f, 0.072169f);
    const f3 bias(0.001f, 0.001f, 0.001f);
    for (int ty = 0; ty < H; ++ty)
        for (int tx = 0; tx < W; ++tx) {
            const int x = tx + vx, y = ty + vy;
            if (x >= W || y >= H) continue;
            f3 Total(0, 0, 0);
            if (U.BFSize > 0) {
                const float sigma = U.BFSigma, sigmaL = U.BFSigmaLum * 100.0f;
                const f3 cc = Dif.rgb(x, y) + bias;
                const f3 cs = Spc.rgb(x, y) + bias;
                const float clum = dot(cc, lumW);
                float Weight = 0;
                // taps: i walks rows, j walks columns (DenoiseJBF:137,147-153)
                for (int i = -5; i <= 5; i += 2)
                    for (int j = -5; j <= 5; j += 2) {
                        const f3 Ci = Dif.rgb(x + j, y + i) + bias;
                        const float lumi = dot(Ci, lumW);
                        const float dist = clampf(float(i * i + j * j) / float(5 * 5), 0.0f, 1.0f);
                        const float dl = (clum - lumi) * (clum - lumi);
                        const float Fi = expf(-dist * dist / (2.0f * sigma * sigma));
                        const float Li = expf(-dl * dl / (2.0f * sigmaL * sigmaL));
                        Total = Total + Ci * Fi * Li;
                        Weight += Fi * Li;
                    }
                Total = Total / Weight;
                if (!U.DebugDraw_Lighting) Total = Total * Alb.rgb(x, y) + cs;
            } else {
                if (U.DebugDraw_Lighting) Total = Dif.rgb(x, y) * f3(0.5f, 0.5f, 0.5f) + Spc.rgb(x, y);
                else Total = Dif.rgb(x, y) * Alb.rgb(x, y) + Spc.rgb(x, y);
            }
            const float eThis = edgeDetect(U.SelectedId, Id0, x, y) ? 0.5f : 0.0f;
            const float ePrev = edgeDetect(U.SelectedId, Id1, x, y) ? 0.5f : 0.0f;
            if (eThis + eThis > 0) Total = mix3(Total, f3(150, 100, 0), eThis + ePrev);
            f3 o;
            if (U.HDR) {
                Total = Total / 2000.f;
                Total = gt3(Total);
                Total = Total * 2000.f;
                o = st2084(Total * U.PaperWhiteNit / 230.0f);
            } else {
                o = gt3(Total * U.PaperWhiteNit / 40000.0f);
            }
            store(out, W, H, x, y, o, 1.0f);
        }
}

} // namespace orc
