"""Prints what Gorp.extract returns for every line of a UTF-8 text file, one TSV row per line, in the same format as
java/JavaParity.java:  <lineNo>\t<MISS | CAPTURE_FAIL | id\tname=value...>

    python -m tests.dump_results definition.grp lines.txt            # CUDA path (needs a GPU)
    python -m tests.dump_results definition.grp lines.txt --oracle   # CPU oracle
"""
import sys


def main():
    definition = open(sys.argv[1], encoding="utf-8").read()
    text = open(sys.argv[2], encoding="utf-8", newline="").read()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    if "--oracle" in sys.argv:
        from oracle import gorp_oracle
        g = gorp_oracle.Gorp(definition)
        for i, ln in enumerate(lines):
            try:
                m = g.extract_map(ln, "\0id")
            except gorp_oracle.ExtractionError:
                print("%d\tCAPTURE_FAIL" % i)
                continue
            if m is None:
                print("%d\tMISS" % i)
            else:
                ident = m.pop("\0id")
                print("\t".join([str(i), ident] + ["%s=%s" % kv for kv in m.items()]))
        return
    from gorp_b200 import DefinitionReader, ExtractionException
    g = DefinitionReader.reader(definition).read()
    b = g.extract_batch_lines(lines)
    for i in range(len(lines)):
        try:
            r = b.result(i)
        except ExtractionException:
            print("%d\tCAPTURE_FAIL" % i)
            continue
        if r is None:
            print("%d\tMISS" % i)
        else:
            print("\t".join([str(i), r.getId()] + ["%s=%s" % kv for kv in r.asMap().items()]))


if __name__ == "__main__":
    main()
