/*
 * Harness for anyone with a JDK and the gorp jar: prints, for every line of a UTF-8 text file, what REAL gorp's
 * Gorp.extract returns, in the same text format as `python -m tests.dump_results` prints for the CUDA path and the
 * oracle. Diffing the two turns "parity unpinned" (restatement of brics + java.util.regex) into "pinned".
 * Not compiled or run in this repository's environment (no JVM there).
 *
 *   javac -cp gorp-core.jar JavaParity.java && java -cp gorp-core.jar:. JavaParity definition.grp lines.txt
 */
import java.io.File;
import java.nio.charset.StandardCharsets;
import java.nio.file.Files;
import java.util.concurrent.*;
import com.salesforce.gorp.*;

public class JavaParity {
    public static void main(String[] a) throws Exception {
        Gorp gorp = DefinitionReader.reader(new File(a[0])).read();
        String text = new String(Files.readAllBytes(new File(a[1]).toPath()), StandardCharsets.UTF_8);
        int start = 0, n = text.length(), lineNo = 0;
        StringBuilder sb = new StringBuilder();
        while (start < n) {
            int nl = text.indexOf('\n', start);
            if (nl < 0) nl = n;
            String line = text.substring(start, nl);
            sb.setLength(0);
            sb.append(lineNo++).append('\t');
            try {
                ExtractionResult r = gorp.extract(line);
                if (r == null) sb.append("MISS");
                else {
                    sb.append(r.getId());
                    for (java.util.Map.Entry<String, Object> e : r.asMap().entrySet())
                        sb.append('\t').append(e.getKey()).append('=').append(e.getValue());
                }
            } catch (ExtractionException e) {
                sb.append("CAPTURE_FAIL");
            }
            System.out.println(sb);
            start = nl + 1;
        }
    }
}
