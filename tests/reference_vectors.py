"""Golden vectors held by the reference's own tests for the hot path.

Each entry cites the reference test it was transcribed from (paths relative to
/root/reference/gorp-core/src/test/java/com/salesforce/gorp/). The Java string literals are
transcribed to their runtime values (e.g. "\\\\s" in Java source == `\\s` here).
Used by tests/test_oracle_golden.py (oracle pin) and tests/test_gpu_parity.py (CUDA path).
"""

# ---- autom/MultiPatternTest.java:12-27 : raw brics-dialect regexes -> accept lists
MULTI_PATTERNS = ["ab+", "abc+", "ab?c", "v", "v.*", "(def)+"]
MULTI_CASES = [
    ("ab", [0]), ("abc", [1, 2]), ("ac", [2]), ("", []), ("v", [3, 4]),
    ("defdef", [5]), ("defde", []), ("abbbbb", [0]),
]

# ---- PolyMatchTest.java : DSL -> DFA accept lists
POLY_SIMPLE = (  # :15-35
    "pattern %word (\\w+)\n"
    "template @base %word\n"
    "extract rule1 {  \n"
    "  template @base value=$value(%word) value2=$value2(%word)\n"
    "}\n"
    "extract rule2 {  \n"
    "  template value=%word\n"
    "}\n",
    [("value=stuff", [1]), ("prefix value=a value2=b", [0])],
)
POLY_INTERMEDIATE = (  # :40-56
    "pattern %phrase \\S+\n"
    "pattern %num \\d+\n"
    "pattern %ts %phrase\n"
    "extract interm {  \n"
    "  template <%num> (foo)[bar] $eventTimeStamp(%ts) end:'$timestamp(%ts)' THE END.\n"
    "}\n",
    [("<123> (foo)[bar] 12:30:58 end:'15:07:00Z' THE END.", [0])],
)
POLY_QUOTED = (  # :61-83
    "pattern %word (\\w+)\n"
    "pattern %quoted \\\"[^\\\"]*\\\"\n"
    "extract quoted {  \n"
    "  template header value=$value(%quoted)\n"
    "}\n"
    "extract unquoted {  \n"
    "  template header value=$value(%word)\n"
    "}\n",
    [("header value=stuff", [1]), ('header value="stuff"', [0])],
)
POLY_COMPLEX = (  # :88-117 (note the literal TAB inside [^ \t]+)
    "pattern %word [a-zA-Z]+\n"
    "pattern %phrase [^ \t]+\n"
    "pattern %num ([0-9]+)\n"
    "pattern %ts %phrase\n"
    "pattern %ip %phrase\n"
    "pattern %maybeUUID %phrase\n"
    "pattern %hostname %phrase\n"
    "template @base <%num>$eventTimeStamp(%ts) $logAgent(%ip) RealSource: \"$logSrcIp(%ip)\"\\\n"
    " Environment: \"$environment(%phrase)\"\\\n"
    " UUID: \"$uuid(%maybeUUID)\"\\\n"
    " RawMsg: <%num>$rawMsgTS(%word %num %phrase) $logSrcHostname(%hostname) $appname(%word)[$appPID(%num)]\n"
    "\n"
    "extract baseMatch {\n"
    "  template @base\n"
    "}\n",
    [('<86>2015-05-12T20:57:53.302858+00:00 10.1.11.141 RealSource: "10.10.5.3"'
      ' Environment: "TEST"'
      ' UUID: "NONE"'
      ' RawMsg: <123>something 1324 keyboard-interactive/pam google.com sshd[137]', [0])],
)
POLY_DSL = [POLY_SIMPLE, POLY_INTERMEDIATE, POLY_QUOTED, POLY_COMPLEX]

# ---- FullExtractionTest.java : full extract (id + captured strings)
FULL_SIMPLE = (  # :13-41
    "pattern %word ([a-zA-Z]+)\n"
    "template @base %word\n"
    "extract double {  \n"
    "  template @base value=$value(%word) value2=$value2(%word)\n"
    "}\n"
    "extract single {  \n"
    "  template value=$value(%word)\n"
    "}\n",
    [("value=foobar", {"id": "single", "value": "foobar"}),
     ("prefix value=a value2=b", {"id": "double", "value": "a", "value2": "b"})],
)
FULL_INTERMEDIATE = (  # :46-65
    "pattern %ws \\s+\n"
    "pattern %word [a-zA-Z]+\n"
    "pattern %phrase \\S+\n"
    "pattern %num \\d+\n"
    "pattern %ts %phrase\n"
    "pattern %ip %phrase\n"
    "extract interm {  \n"
    "  template <%num>$eventTimeStamp(%ts) $logAgent(%ip) RealSource: \"$logSrcIp(%ip)\"\n"
    "}\n",
    [('<86>2015-05-12T20:57:53.302858+00:00 10.1.11.141 RealSource: "10.10.5.3"',
      {"id": "interm", "eventTimeStamp": "2015-05-12T20:57:53.302858+00:00", "logAgent": "10.1.11.141",
       "logSrcIp": "10.10.5.3"})],
)
_FULL_DEF = (  # :70-107, after DEF.replace('\'', '"')
    "### First, let's define basic patterns using 'patterns' (regexps)\n"
    "# 'phrase' means non-space-sequence of characters; 'word' letters; 'num' digits\n"
    "pattern %word [a-zA-Z]+\n"
    "pattern %phrase \\S+\n"
    "pattern %num \\d+\n"
    "# more semantic macros, loosely defined\n"
    "pattern %ts %phrase\n"
    "pattern %ip %phrase\n"
    "pattern %maybeUUID %phrase\n"
    "pattern %hostname %phrase\n"
    "pattern %any .*\n"
    "\n"
    "# then possible 'templates', building blocks that consist of named patterns, literal text and possible embedded\n"
    "# 'anonymous' patterns (enclosed in %{....} and neither parsed (to substituted) nor escaped (like literal text))\n"
    "\n"
    "template @base <%num>$eventTimeStamp(%ts) $logAgent(%ip) RealSource: '$logSrcIp(%ip)'\\\n"
    " Environment: '$environment(%phrase)' UUID: '$uuid(%maybeUUID)'\\\n"
    " RawMsg: <%num>$rawMsgTS(%word %num %phrase) $logSrcHostname(%hostname)\\\n"
    " $appname(%word)[$appPID(%num)]\n"
    "\n"
    "# and then higher-level composition\n"
    "\n"
    "# sample:\n"
    "#<86>2015-03-16T20:57:53.302858+00:00 10.1.11.141 RealSource: '10.1.2.72' Environment: 'TEST' UUID: 'NO'\n"
    "# RawMsg: <86>Apr 16 20:54:53 host-prodnet sshd[12973]: Accepted keyboard-interactive/pam for badguy.ru from 1.2.3.4 port 58216 ssh2\n"
    "\n"
    "extract sshdMatch {\n"
    "  template @base: $authStatus(Accepted) $sshAuthMethod(%phrase) for $user(%hostname)\\\n"
    " from $srcIP(%ip) port $srcPort(%num) $sshProtocol(%phrase)\n"
    "  append 'service':'ssh', 'logType':'security', 'serviceType':'authentication' \n"
    "}\n"
    "extract baseMatch {\n"
    "  template @base\n"
    "}\n"
).replace("'", '"')
_FULL_IN1 = ("<86>2015-05-12T20:57:53.302858+00:00 10.1.11.141 RealSource:   '10.10.5.3'"
             " Environment: 'TEST' UUID: 'NO'"
             " RawMsg: <123>something 1324 more-or-less google.com sshd[137]").replace("'", '"')
_FULL_IN2 = (_FULL_IN1 + ": Accepted keyboard-interactive/pam for badguy.ru from 1.2.3.4 port 58216 ssh2")
FULL_FULL = (
    _FULL_DEF,
    [(_FULL_IN1, {"id": "baseMatch", "logSrcIp": "10.10.5.3", "environment": "TEST", "uuid": "NO",
                  "rawMsgTS": "something 1324 more-or-less", "logSrcHostname": "google.com",
                  "appname": "sshd", "appPID": "137"}),
     (_FULL_IN2, {"id": "sshdMatch", "user": "badguy.ru", "sshProtocol": "ssh2", "srcIP": "1.2.3.4",
                  "srcPort": "58216", "authStatus": "Accepted", "sshAuthMethod": "keyboard-interactive/pam",
                  "service": "ssh", "logType": "security", "serviceType": "authentication"})],
)
PARAM_EXTRACTOR = (  # ParametricExtractorTest.java:14-33
    "pattern %num ([0-9]+)\n"
    "pattern %word ([a-zA-Z]+)\n"
    "pattern %ip [a-zA-Z\\.]+\n"
    "template @ip %ip\n"
    "template @port %num\n"
    "template @endpoint() $1(@ip):$2(@port)\n"
    "extract Net {  \n"
    "  template @endpoint($srcIp,$srcPort)/%word\n"
    "}\n",
    [("foo.bar.com:8080/user", {"id": "Net", "srcIp": "foo.bar.com", "srcPort": "8080"})],
)
PARAM_TEMPLATE = (  # ParametricTemplateTest.java:14-37
    "pattern %word ([a-zA-Z]+)\n"
    "pattern %num ([0-9]+)\n"
    "pattern %ip [a-zA-Z\\.]+\n"
    "template @ip %ip\n"
    "template @port %num\n"
    "template @colonPair() @1:@2\n"
    "extract Net {  \n"
    "  template $endpoint(@colonPair(@ip,@port))/%word\n"
    "}\n",
    [("foo.bar.com:8080/user", {"id": "Net", "endpoint": "foo.bar.com:8080"})],
)
# exact map sizes asserted by the reference (asMap("id") / asMap())
FULL_EXACT = [FULL_SIMPLE, PARAM_EXTRACTOR, PARAM_TEMPLATE]
# the reference asserts only a subset of keys for these (the rest is derived)
FULL_SUBSET = [FULL_INTERMEDIATE, FULL_FULL]

# ---- util/RegexHelperTest.java:10-37 : translator known answers
CC_d, CC_s, CC_w = "0-9", " \b\f\n\r\t", "a-zA-Z_0-9"
QUOTE_KATS = [("", ""), ("(foo)", "\\(foo\\)"), ("[foo]", "\\[foo\\]"), ("a\\b", "a\\\\b")]
AUTOM_KATS = [("", ""), ("[\\w]+", "[" + CC_w + "]+"), ("\\w+", "[" + CC_w + "]+"),
              ("[\\d\\s]+", "[" + CC_d + CC_s + "]+")]
JDK_KATS = [("", ""), ("stuff([ab]+([de]+))", "stuff(?:[ab]+(?:[de]+))"), ("stuff\\(sic\\)", "stuff\\(sic\\)")]

# ---- SURVEY Appendix I : hand-derived generated strings (\t == literal TAB)
_S = "[^ \b\f\n\r\t]"
DERIVED_STRINGS = [
    (FULL_SIMPLE[0], [
        ("double", "([a-zA-Z]+)[ \t]+value=(([a-zA-Z]+))[ \t]+value2=(([a-zA-Z]+))",
         "(?:[a-zA-Z]+)[ \t]+value=((?:[a-zA-Z]+))[ \t]+value2=((?:[a-zA-Z]+))", ["value", "value2"]),
        ("single", "value=(([a-zA-Z]+))", "value=((?:[a-zA-Z]+))", ["value"])]),
    (POLY_SIMPLE[0], [
        ("rule1", "([a-zA-Z_0-9]+)[ \t]+value=(([a-zA-Z_0-9]+))[ \t]+value2=(([a-zA-Z_0-9]+))",
         "(?:\\w+)[ \t]+value=((?:\\w+))[ \t]+value2=((?:\\w+))", ["value", "value2"]),
        ("rule2", "value=([a-zA-Z_0-9]+)", "value=(?:\\w+)", [])]),
    (POLY_QUOTED[0], [
        ("quoted", "header[ \t]+value=(\\\"[^\\\"]*\\\")", "header[ \t]+value=(\\\"[^\\\"]*\\\")", ["value"]),
        ("unquoted", "header[ \t]+value=(([a-zA-Z_0-9]+))", "header[ \t]+value=((?:\\w+))", ["value"])]),
    (POLY_INTERMEDIATE[0], [
        ("interm",
         "\\<[0-9]+\\>[ \t]+\\(foo\\)\\[bar\\][ \t]+(" + _S + "+)[ \t]+end:'(" + _S + "+)'[ \t]+THE[ \t]+END\\.",
         "\\<\\d+\\>[ \t]+\\(foo\\)\\[bar\\][ \t]+(\\S+)[ \t]+end:'(\\S+)'[ \t]+THE[ \t]+END\\.",
         ["eventTimeStamp", "timestamp"])]),
    (PARAM_EXTRACTOR[0], [
        ("Net", "([a-zA-Z\\.]+):(([0-9]+))/([a-zA-Z]+)", "([a-zA-Z\\.]+):((?:[0-9]+))/(?:[a-zA-Z]+)",
         ["srcIp", "srcPort"])]),
    (PARAM_TEMPLATE[0], [
        ("Net", "([a-zA-Z\\.]+:([0-9]+))/([a-zA-Z]+)", "([a-zA-Z\\.]+:(?:[0-9]+))/(?:[a-zA-Z]+)", ["endpoint"])]),
]

SIMPLE_GRP = (  # /root/reference/samples/simple.grp:4-23 (config #1); note the trailing space in the template
    "pattern %ws \\s+\n"
    "pattern %optws \\s*\n"
    "pattern %word \\w+\n"
    "pattern %phrase \\S+\n"
    "pattern %num \\d+\n"
    "pattern %ts %phrase\n"
    "pattern %ip %phrase\n"
    "pattern %any .*\n"
    "template @base <%num>$eventTimeStamp(%ts)\n"
    "extract sampleMatch {\n"
    "  template @base ($authStatus(Accepted)) \n"
    "}\n"
)
README_DEF = (  # /root/reference/README.md:115-134 (config #2)
    "pattern %num \\d+\n"
    "pattern %word \\w+\n"
    "pattern %phrase \\S+\n"
    "\n"
    "extract PutRequest {\n"
    "   # It's ok to: (a) extract constant value; (b) concatenate physical lines with backslash\n"
    "   template [$timestamp(%num)]: $verb(PUT) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
    "extract GetRequest {\n"
    "   template [$timestamp(%num)]: $verb(GET) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
    "extract OtherRequest {\n"
    "   template [$timestamp(%num)]: $verb(%word) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
)

# ---- definition-error known answers (message substrings, case-insensitive; TestBase.verifyException)
ERROR_KATS = [
    # ParametricTemplateTest.java:42-154
    ("pattern %word ([a-zA-Z]+)\ntemplate @pair() @1:@2\ntemplate @full @pair\nextract Result {  \n  template @full\n}\n",
     ["Missing parameter list", "@pair"]),
    ("template @pair @1:@2\ntemplate @full xyz\nextract Result {  \n  template @full\n}\n",
     ["Invalid variable reference"]),
    ("template @pair() @1:@2\ntemplate @a    a\ntemplate @full @pair(@a\nextract Result {  \n  template @full\n}\n",
     ["Unexpected end of line"]),
    ("template @constant text\ntemplate @abc @full(@ab(@c,@1))\ntemplate @full() @1\nextract Result {  \n  template @full\n}\n",
     ["non-existing template '@ab'"]),
    ("template @pair() @1:@2\ntemplate @foo foosball\ntemplate @fooPair @pair(@foo)\nextract Result {  \n  template @fooPair\n}\n",
     ["Parameter mismatch"]),
    ("template @pair() @1:@2\ntemplate @foo foosball\ntemplate @fooPair @pair(@foo,@foo,@foo)\nextract Result {  \n  template @fooPair\n}\n",
     ["Parameter mismatch"]),
    # ParametricExtractorTest.java:42-62
    ("pattern %num ([0-9]+)\npattern %word ([a-zA-Z]+)\npattern %ip [a-zA-Z\\.]+\ntemplate @ip %ip\ntemplate @port %num\n"
     "template @endpoint() $1(@ip):$2(@port)\nextract Net {  \n  template @endpoint($srcIp,$srcPort)/%word @endpoint($srcIp,$whatever)\n}\n",
     ["duplicate extractor name", "srcIp"]),
    # ExtractionResolutionTest.java:68-95
    ("pattern %a a\ntemplate @base (%a:foo)\n", ["No extraction definitions found"]),
    ("pattern %word \\w+\ntemplate @extr $value(%word)\nextract match {  \n  template @extr @extr\n}\n",
     ["Duplicate extractor name"]),
    # UncookedDefTest.java:215-243
    ("pattern %'ws' \\s+\npattern %optws \\s*\npattern %ws \\S+\n", ["duplicate"]),
    ("pattern %'ws' \\s+%\n", ["Orphan '%'"]),
    # PatternResolutionTest.java:40-68 (+ a trailing extraction so that read() reaches resolution)
    ("pattern %a Ok: %b\npattern %b But... %c\nextract x {\n template %a\n}\n", ["non-existing pattern '%c'"]),
    ("pattern %a Kaboom: %a\nextract x {\n template %a\n}\n", ["cyclic pattern reference to '%a'"]),
    ("pattern %a %b\npattern %b %a\nextract x {\n template %a\n}\n", ["cyclic pattern reference to '%a'"]),
    # io/InputLineReaderTest.java:38-54
    ("pattern %a a\npattern %b b\npattern %c combo... \\", ["unexpected end-of-input when expecting line continuation", "row 3"]),
]
