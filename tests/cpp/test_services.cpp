// C++ consumer of the drop-in boundary: exercises include/sa_services.hpp against libsa_engine.so the way
// the reference's Java services would, with known-answer checks (impulse, on-bin tone, error mapping).
#include <cmath>
#include <cstdio>
#include <complex>
#include <vector>

#include "sa_services.hpp"

using namespace spectral_analyzer;

#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    Engine eng(0);
    SpectralService svc(eng);
    const int n = 1024;
    // on-bin tone k = 37, cf32_le
    std::vector<float> iq(2 * 4 * n);
    for (int i = 0; i < 4 * n; i++) {
        const double ph = 2.0 * M_PI * 37.0 * (i % n) / n;
        iq[2 * i] = (float)std::cos(ph); iq[2 * i + 1] = (float)std::sin(ph);
    }
    MappedByteBuffer buf{iq.data(), iq.size() * sizeof(float)};
    std::vector<double> mag = svc.computeMagnitudes(buf, 0, n, "cf32_le");
    int arg = 0;
    for (int i = 1; i < n; i++) if (mag[i] > mag[arg]) arg = i;
    EXPECT(arg == (37 + n / 2) % n);                               // fft-shifted, SpectralService.java:78
    EXPECT(std::fabs(mag[arg] - 20.0 * std::log10((double)n)) < 1e-3);
    // waterfall: 5 columns requested, 4 available -> last row is -150 (MainController.java:994-998)
    std::vector<double> wf = svc.computeWaterfall(buf, 0, 5, n, "cf32_le");
    EXPECT(wf[4 * n + 3] == -150.0 && std::fabs(wf[3 * n + arg] - mag[arg]) < 1e-3);
    // error mapping
    bool threw = false;
    try { svc.computeMagnitudes(buf, 0, 1000, "cf32_le"); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { svc.computeMagnitudes(buf, (int)buf.capacity - 8, n, "cf32_le"); } catch (const std::out_of_range&) { threw = true; }
    EXPECT(threw);
    // downconvert the tone to DC and take its PSD
    ExtractDownConvertService dc(eng);
    auto z = dc.extractAndDownConvert(buf, 0, 4 * n, "cf32_le", 37.0 / n, 4);
    EXPECT(z[0].size() == (size_t)n);
    EXPECT(std::fabs(z[0][n - 1] - 1.0) < 1e-4 && std::fabs(z[1][n - 1]) < 1e-4);
    auto psd = PowerSpectralDensity::calculatePsdWelch(eng, z, 1.0e6 / 4, 256);
    int pk = 0;
    for (int i = 1; i < 256; i++) if (psd[1][i] > psd[1][pk]) pk = i;
    EXPECT(std::fabs(psd[0][pk]) < 1.0e6 / 4 / 256 * 1.5);
    // IqData.getInterleavedBinary: DC tone 1.0 -> int16 32767 / 0 (truncation), float32 little-endian
    IqData iqd(eng, z);
    std::vector<uint8_t> i16 = iqd.getInterleavedBinary("int16");
    EXPECT(i16.size() == 4 * (size_t)n);
    const int16_t last_i = (int16_t)(i16[4 * (n - 1)] | (i16[4 * (n - 1) + 1] << 8));
    EXPECT(last_i >= 32763 && last_i <= 32767);
    threw = false;
    try { iqd.getInterleavedBinary("int8"); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw);
    // renderSpectrogram: 4 columns x 64 rows; the tone's row is the brightest of column 0
    std::vector<uint8_t> px = SpectrogramRenderer::renderSpectrogram(eng, buf, 0, 4, 64, n, "cf32_le", 1.0e6, -160.0, -30.0,
                                                                    SA_CMAP_GRAYSCALE, 1, SA_REDUCE_MAX);
    EXPECT(px.size() == 4u * 64u * 4u);
    int best = 0;
    for (int y = 1; y < 64; y++) if (px[(size_t)(y * 4) * 4] > px[(size_t)(best * 4) * 4]) best = y;
    const int f_row = 63 - best;                                   // y flipped, MainController.java:1288
    EXPECT((int)((double)f_row / 64 * n) <= arg && arg < (int)((double)(f_row + 1) / 64 * n));
    // tile cache: a 64-pt view over the same buffer, scrolled by 3 columns, equals the direct render; tiles are reused
    {
        const int nf = 64, W = 40, H = 16;
        CanvasTileCache cache(eng, 16, 8);
        for (int64_t start : {0, 3 * nf, 3 * nf + 5, 19 * nf + 5}) {
            std::vector<uint8_t> direct = SpectrogramRenderer::renderSpectrogram(eng, buf, start, W, H, nf, "cf32_le", 1.0e6,
                                                                                -160.0, -30.0, SA_CMAP_HEATMAP);
            std::vector<uint8_t> tiled = cache.view(buf, start, W, H, nf, "cf32_le", 1.0e6, -160.0, -30.0, SA_CMAP_HEATMAP);
            EXPECT(direct == tiled);
        }
        EXPECT(cache.hits() > 0 && cache.misses() > 0);
    }
    // the annotation batch of AnnotationController.java:321-360 as one call: rows equal the per-annotation service,
    // PSD rows equal calculatePsdWelch of each row; a short annotation follows the single-window rule (rest NaN)
    {
        std::vector<ExtractDownConvertService::Annotation> rows = { {0, 4 * n, 37.0 / n, 4}, {512, 2048, -0.1, 8}, {100, 900, 0.2, 4} };
        std::vector<double> psd_rows;
        auto batch = dc.extractAndDownConvertBatch(buf, "cf32_le", 1.0e6, rows, 256, &psd_rows);
        EXPECT(batch.size() == rows.size() && psd_rows.size() == rows.size() * 256);
        for (size_t i = 0; i < rows.size(); i++) {
            auto one = dc.extractAndDownConvert(buf, rows[i].startSample, rows[i].count, "cf32_le", rows[i].freqOff, rows[i].down);
            EXPECT(batch[i][0] == one[0] && batch[i][1] == one[1]);
            const int m = (int)one[0].size(), nf = m < 256 ? m : 256;
            auto p1 = PowerSpectralDensity::calculatePsdWelch(eng, one, 1.0e6 / rows[i].down, nf);
            for (int k = 0; k < nf; k++) EXPECT(std::fabs(psd_rows[i * 256 + k] - p1[1][k]) < 1e-9);
            for (int k = nf; k < 256; k++) EXPECT(std::isnan(psd_rows[i * 256 + k]));
        }
        // a JDSP profile with delay compensation and ceil length: one more output, the DC tone settles 4*down samples earlier
        sa_analysis_config cfg;
        sa_analysis_config_init(&cfg);
        cfg.delay_mode = SA_DELAY_SAME; cfg.length_mode = SA_LEN_CEIL;
        eng.setAnalysisConfig(cfg);
        auto zs = dc.extractAndDownConvert(buf, 0, 4 * n - 3, "cf32_le", 37.0 / n, 4);
        EXPECT(zs[0].size() == (size_t)n);                              // ceil((4n - 3) / 4)
        EXPECT(std::fabs(zs[0][4] - 1.0) < 1e-3 && std::fabs(z[0][4] - 1.0) > 1e-3);    // all 33 taps overlap data at m = 4 only when centred
        eng.resetAnalysisConfig();
    }
    std::printf("cpp services ok\n");
    return 0;
}
