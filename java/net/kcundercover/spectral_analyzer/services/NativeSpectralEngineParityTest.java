package net.kcundercover.spectral_analyzer.services;

import static org.junit.jupiter.api.Assertions.assertEquals;
import static org.junit.jupiter.api.Assertions.assertTrue;

import java.nio.ByteOrder;
import java.nio.MappedByteBuffer;
import java.nio.channels.FileChannel;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.StandardOpenOption;
import java.util.Random;

import org.junit.jupiter.api.Test;
import org.junit.jupiter.api.io.TempDir;

/**
 * The headless parity + timing test BASELINE.json config 1 names ("headless spectrogram of a synthetic cf32
 * SigMF recording, 10 M samples, via ./gradlew test"): the reference has no numeric test today
 * (src/test/.../SpectralAnalyzerApplicationTests.java is a Spring contextLoads), so this is the JUnit 5 test a
 * maintainer adds next to NativeSpectralEngine.  It goes under src/test/java/.../services/ of the reference.
 *
 * NOT compiled in the engine repository (no JVM in the build image); the same comparison runs there against the
 * C / numpy restatement of SpectralService.computeMagnitudes (tests/test_gpu_spectrogram.py).
 *
 * Parity mode = the reference's own framing: rectangular window, hop = nfft, 20*log10(|X| + 1e-10), fft-shifted
 * (SpectralService.java:33-85, MainController.java:982-999).  Tolerance of BASELINE.json: power within 1e-5
 * relative on signal bins (4.3e-5 dB), 1e-3 dB on bins above the noise floor.
 */
class NativeSpectralEngineParityTest {
    private static final int N_SAMPLES = 10_000_000;
    private static final int NFFT = 1024;

    private static MappedByteBuffer syntheticCf32(Path file) throws Exception {
        try (FileChannel ch = FileChannel.open(file, StandardOpenOption.CREATE, StandardOpenOption.READ, StandardOpenOption.WRITE)) {
            MappedByteBuffer w = ch.map(FileChannel.MapMode.READ_WRITE, 0, 8L * N_SAMPLES);
            w.order(ByteOrder.LITTLE_ENDIAN);
            Random rng = new Random(1);
            double[][] tones = {{0.1250, 0.5}, {-0.28137, 0.25}, {0.40213, 0.125}};   // cycles/sample, amplitude
            for (int n = 0; n < N_SAMPLES; n++) {
                double re = 0.005 * rng.nextGaussian(), im = 0.005 * rng.nextGaussian();
                for (double[] t : tones) {
                    double ph = 2 * Math.PI * ((t[0] * n) % 1.0);
                    re += t[1] * Math.cos(ph);
                    im += t[1] * Math.sin(ph);
                }
                w.putFloat((float) re);
                w.putFloat((float) im);
            }
            w.force();
            MappedByteBuffer r = ch.map(FileChannel.MapMode.READ_ONLY, 0, 8L * N_SAMPLES);
            r.order(ByteOrder.LITTLE_ENDIAN);                                        // SigMfHelper.java:87-91
            return r;
        }
    }

    @Test
    void batchedEngineMatchesComputeMagnitudes(@TempDir Path tmp) throws Throwable {
        MappedByteBuffer buf = syntheticCf32(tmp.resolve("synthetic.sigmf-data"));
        int frames = N_SAMPLES / NFFT;
        SpectralService reference = new SpectralService();                           // commons-math3 FP64 path
        try (NativeSpectralEngine engine = new NativeSpectralEngine(0)) {
            long t0 = System.nanoTime();
            double[][] got = engine.computeWaterfall(buf, 0L, frames, NFFT, "cf32_le");   // ONE downcall
            long tEngine = System.nanoTime() - t0;
            t0 = System.nanoTime();
            double worstStrong = 0, worstAbove = 0;
            for (int t = 0; t < frames; t++) {
                double[] ref = reference.computeMagnitudes(buf, t * NFFT * 8, NFFT, "cf32_le");   // :33
                double max = Double.NEGATIVE_INFINITY;
                double[] sorted = ref.clone();
                java.util.Arrays.sort(sorted);
                double floor = sorted[NFFT / 2];                                     // median bin = noise floor
                for (double v : ref) max = Math.max(max, v);
                for (int k = 0; k < NFFT; k++) {
                    double d = Math.abs(got[t][k] - ref[k]);
                    if (ref[k] >= max - 40) worstStrong = Math.max(worstStrong, d);
                    if (ref[k] >= floor + 10) worstAbove = Math.max(worstAbove, d);
                }
            }
            long tReference = System.nanoTime() - t0;
            assertEquals(frames, got.length);
            assertTrue(worstStrong <= 4.35e-5, "signal bins differ by " + worstStrong + " dB");
            assertTrue(worstAbove <= 1e-3, "bins above the floor differ by " + worstAbove + " dB");
            System.out.printf("spectrogram of %d samples: engine %.1f ms, reference loop %.1f ms (%.1fx)%n",
                    N_SAMPLES, tEngine / 1e6, tReference / 1e6, (double) tReference / tEngine);
        }
    }

    @Test
    void errorsMapToTheExceptionsTheCallersExpect(@TempDir Path tmp) throws Throwable {
        MappedByteBuffer buf = syntheticCf32(tmp.resolve("synthetic.sigmf-data"));
        try (NativeSpectralEngine engine = new NativeSpectralEngine(0)) {
            org.junit.jupiter.api.Assertions.assertThrows(IllegalArgumentException.class,
                    () -> engine.computeMagnitudes(buf, 0, 1000, "cf32_le"));        // not a power of two
            org.junit.jupiter.api.Assertions.assertThrows(IndexOutOfBoundsException.class,
                    () -> engine.computeMagnitudes(buf, buf.capacity() - 8, NFFT, "cf32_le"));
        }
    }
}
