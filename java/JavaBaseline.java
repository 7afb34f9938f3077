/*
 * The true JVM baseline: Gorp.extract over an ExecutorService, all cores, lines/s. Not run here (no JVM).
 *   java -cp gorp-core.jar:. JavaBaseline definition.grp lines.txt [threads]
 */
import java.io.File;
import java.nio.charset.StandardCharsets;
import java.nio.file.Files;
import java.util.*;
import java.util.concurrent.*;
import com.salesforce.gorp.*;

public class JavaBaseline {
    public static void main(String[] a) throws Exception {
        final Gorp gorp = DefinitionReader.reader(new File(a[0])).read();
        final List<String> lines = Files.readAllLines(new File(a[1]).toPath(), StandardCharsets.UTF_8);
        final int threads = a.length > 2 ? Integer.parseInt(a[2]) : Runtime.getRuntime().availableProcessors();
        ExecutorService pool = Executors.newFixedThreadPool(threads);
        for (int round = 0; round < 5; ++round) {
            long t0 = System.nanoTime();
            List<Future<long[]>> fs = new ArrayList<>();
            for (int t = 0; t < threads; ++t) {
                final int lo = (int) ((long) lines.size() * t / threads), hi = (int) ((long) lines.size() * (t + 1) / threads);
                fs.add(pool.submit(() -> {
                    long hit = 0, fail = 0;
                    for (int i = lo; i < hi; ++i) {
                        try { if (gorp.extract(lines.get(i)) != null) ++hit; } catch (ExtractionException e) { ++fail; }
                    }
                    return new long[] {hit, fail};
                }));
            }
            long hit = 0, fail = 0;
            for (Future<long[]> f : fs) { long[] r = f.get(); hit += r[0]; fail += r[1]; }
            double s = (System.nanoTime() - t0) / 1e9;
            System.out.printf("round %d: %d lines, %d matched, %d capture failures, %d threads, %.0f lines/s%n",
                    round, lines.size(), hit, fail, threads, lines.size() / s);
        }
        pool.shutdown();
    }
}
