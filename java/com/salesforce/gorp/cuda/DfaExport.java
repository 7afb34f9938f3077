/*
 * Host-side export of a compiled gorp definition for libgorpcuda (JDK 22+; NOT compiled or tested in the build
 * environment of this repository: no JVM exists there — see INTEGRATION.md).
 *
 * Serialises exactly what the reference holds after Gorp.construct():
 *   Automata._alphabet / _transitions / _stride / _accept   (autom/Automata.java:23-26, private => reflection,
 *                                                             the same way gorp itself reaches into brics,
 *                                                             autom/DkBricsAutomatonAccess.java:29-34)
 *   CookedExtraction.getName() / getRegexpSource()           (model/CookedExtraction.java:38,46)
 * into the little-endian blob documented in gorp_b200/csrc/host/model.hpp ("GORPDFA1").
 */
package com.salesforce.gorp.cuda;

import java.lang.reflect.Field;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.List;

import com.salesforce.gorp.Gorp;
import com.salesforce.gorp.autom.PolyMatcher;
import com.salesforce.gorp.model.CookedExtraction;

public final class DfaExport {
    private DfaExport() { }

    private static Object field(Object o, String name) throws ReflectiveOperationException {
        Field f = o.getClass().getDeclaredField(name);
        f.setAccessible(true);
        return f.get(o);
    }

    public static byte[] export(Gorp gorp) throws ReflectiveOperationException {
        PolyMatcher matcher = gorp.getMatcher();
        Object automata = field(matcher, "automata");
        int[][] accept = (int[][]) field(automata, "_accept");
        int stride = (Integer) field(automata, "_stride");
        int[] transitions = (int[]) field(automata, "_transitions");
        int[] alphabet = (int[]) field(automata, "_alphabet");
        List<CookedExtraction> extractions = gorp.getExtractions();
        final int nStates = accept.length;

        int acceptTotal = 0;
        for (int[] a : accept) acceptTotal += a.length;
        int size = 40 + 65536 * 2 + 4 * transitions.length + 4 * nStates + 4 * (nStates + 1) + 4 * acceptTotal;
        for (CookedExtraction x : extractions) {
            String[] names = extractorNames(x);
            size += 4 + str(x.getName()) + str("") + str(x.getRegexpSource()) + 4;
            for (String n : names) size += str(n);
            size += 4; // append JSON is applied on the Java side (ExtractionResult.asMap), not shipped
        }
        ByteBuffer b = ByteBuffer.allocate(size).order(ByteOrder.LITTLE_ENDIAN);
        b.put(new byte[] {'G', 'O', 'R', 'P', 'D', 'F', 'A', '1'});
        b.putInt(1).putInt(nStates).putInt(stride).putInt(extractions.size()).putInt(0).putInt(0).putLong(0L);
        for (int c = 0; c < 65536; ++c) b.putShort((short) alphabet[c]);
        for (int t : transitions) b.putInt(t);
        for (int[] a : accept) b.putInt(a.length == 0 ? -1 : a[0]);
        int run = 0;
        b.putInt(0);
        for (int[] a : accept) { run += a.length; b.putInt(run); }
        for (int[] a : accept) for (int v : a) b.putInt(v);
        for (CookedExtraction x : extractions) {
            String[] names = extractorNames(x);
            b.putInt(names.length);
            putStr(b, x.getName());
            putStr(b, ""); // the automaton-dialect string is not retained by Gorp; the tables above carry its language
            putStr(b, x.getRegexpSource());
            b.putInt(names.length);
            for (String n : names) putStr(b, n);
            b.putInt(0);
        }
        byte[] out = b.array();
        long h = 0xcbf29ce484222325L; // FNV-1a 64 over everything after the 40-byte header
        for (int i = 40; i < out.length; ++i) { h ^= (out[i] & 0xff); h *= 0x100000001b3L; }
        ByteBuffer.wrap(out).order(ByteOrder.LITTLE_ENDIAN).putLong(32, h);
        return out;
    }

    private static String[] extractorNames(CookedExtraction x) throws ReflectiveOperationException {
        Field f = CookedExtraction.class.getDeclaredField("_extractorNames"); // protected final
        f.setAccessible(true);
        return (String[]) f.get(x);
    }

    private static int str(String s) { return 4 + ((2 * s.length() + 3) & ~3); }

    private static void putStr(ByteBuffer b, String s) {
        b.putInt(s.length());
        for (int i = 0; i < s.length(); ++i) b.putChar(s.charAt(i));
        while ((b.position() & 3) != 0) b.put((byte) 0);
    }
}
