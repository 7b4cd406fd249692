package net.kcundercover.spectral_analyzer.services;

import java.nio.MappedByteBuffer;
import java.util.LinkedHashMap;
import java.util.Map;

/**
 * Scrolling views on top of {@link NativeSpectralEngine#renderSpectrogram}: the reference redraws every column on
 * each scroll-bar move (MainController.java:319 sets currentSampleOffset, :980-999 recomputes canvasW frames,
 * :1261-1291 repaints).  Here the view is assembled from fixed-width tiles of canvas columns kept in an LRU map, so
 * a scroll step renders only the tiles that enter the view.  Same scheme as spectral_analyzer_b200/tiles.py and
 * CanvasTileCache in include/sa_services.hpp; source only in this repository (no JDK in the build image).
 *
 * Tiles are aligned on global column indices of the view's phase (currentSampleOffset % fftSize); every column's
 * pixels depend on that column's samples only, so the result equals one renderSpectrogram call at the same offset.
 */
public final class CanvasTileCache {
    private final NativeSpectralEngine engine;
    private final int tileW;
    private final LinkedHashMap<String, byte[]> tiles;
    private long hits, misses;

    public CanvasTileCache(NativeSpectralEngine engine, int tileW, final int maxTiles) {
        if (tileW < 1 || maxTiles < 1) throw new IllegalArgumentException("tileW and maxTiles must be positive");
        this.engine = engine;
        this.tileW = tileW;
        this.tiles = new LinkedHashMap<>(16, 0.75f, true) {          // access order = LRU
            @Override protected boolean removeEldestEntry(Map.Entry<String, byte[]> e) { return size() > maxTiles; }
        };
    }

    public long hits() { return hits; }
    public long misses() { return misses; }
    /** Call when SigMfHelper.load maps another file (the buffer identity is part of the key, its content is not). */
    public void clear() { tiles.clear(); }

    /** RGBA8 [canvasH][canvasW], row 0 = top, for the view whose first column starts at currentSampleOffset. */
    public byte[] view(MappedByteBuffer buffer, long currentSampleOffset, int canvasW, int canvasH, int fftSize,
                       String datatype, double sampleRate, double minDb, double maxDb, int colormap) throws Throwable {
        if (currentSampleOffset < 0) throw new IllegalArgumentException("currentSampleOffset must be non-negative");
        final long spc = fftSize;                                     // samples per column (one frame per column)
        final long phase = currentSampleOffset % spc, g0 = currentSampleOffset / spc;
        final byte[] out = new byte[canvasW * canvasH * 4];
        for (long k = g0 / tileW; k <= (g0 + canvasW - 1) / tileW; k++) {
            final String key = System.identityHashCode(buffer) + "/" + buffer.capacity() + "/" + datatype + "/" + fftSize
                    + "/" + canvasH + "/" + colormap + "/" + minDb + "/" + maxDb + "/" + sampleRate + "/" + phase + "/" + k;
            byte[] tile = tiles.get(key);
            if (tile == null) {
                misses++;
                tile = engine.renderSpectrogram(buffer, phase + k * tileW * spc, tileW, canvasH, fftSize, datatype,
                                                sampleRate, minDb, maxDb, colormap);
                tiles.put(key, tile);
            } else {
                hits++;
            }
            final int lo = (int) (Math.max(g0, k * tileW) - g0), hi = (int) (Math.min(g0 + canvasW, (k + 1) * tileW) - g0);
            final int src0 = (int) (g0 + lo - k * tileW);
            for (int y = 0; y < canvasH; y++)
                System.arraycopy(tile, (y * tileW + src0) * 4, out, (y * canvasW + lo) * 4, (hi - lo) * 4);
        }
        return out;
    }
}
