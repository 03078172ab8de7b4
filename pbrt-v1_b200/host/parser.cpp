// parser.cpp -- hand-written lexer + recursive-descent parser for pbrt-v1
// scene files.
//
// The reference generates its front end with flex/bison (core/pbrtlex.l,
// core/pbrtparse.y, core/parser.cpp:27-51); neither tool exists in this
// toolchain, so the grammar is restated here by hand.  The file only uses the
// public surface of "api.h" / "paramset.h" / "pbrt.h", so the very same source
// is also compiled against the reference's headers when oracle/build_ref.sh
// links the reference renderer (oracle/_ref) -- both renderers therefore see
// identical API call sequences for a given scene file.
//
// Token rules: pbrtlex.l:95-176.  Statement forms: pbrtparse.y:294-468.
// Parameter typing ("float foo" -> ParamSet::AddFloat): pbrtparse.y:470-574.
#if __cplusplus >= 201703L
#include <charconv>
#include <system_error>
#endif
#include "api.h"
#include "pbrt.h"
#include "paramset.h"
#include <stdarg.h>
#include <ctype.h>

int line_num = 0;
string current_file;

namespace {

enum TokKind { T_EOF, T_KEYWORD, T_NUM, T_STRING, T_LBRACK, T_RBRACK };

struct Token {
  TokKind kind;
  string text;   // keyword / identifier / string payload
  float num;
};

struct Source {
  FILE *fp;
  string filename;
  int lineNum;
};

void ParseError(const char *format, ...) {
  char msg[4096];
  va_list args;
  va_start(args, format);
  vsnprintf(msg, sizeof(msg), format, args);
  va_end(args);
  Severe("Parsing error: %s", msg);   // pbrtparse.y:36-38 (yyerror)
}

class Lexer {
 public:
  explicit Lexer(FILE *fp) : cur(fp), havePeek(false) {}
  Token Next() {
    if (havePeek) { havePeek = false; return peeked; }
    return Scan();
  }
  const Token &Peek() {
    if (!havePeek) { peeked = Scan(); havePeek = true; }
    return peeked;
  }
 private:
  int Get() {
    if (!pushback.empty()) { int c = pushback.back(); pushback.pop_back(); return c; }
    for (;;) {
      int c = getc_unlocked(cur);   // (one reader; fgetc's lock per character was a third of the parse)
      if (c != EOF) return c;
      if (stack.empty()) return EOF;
      // pbrtlex.l:180-186 (yywrap): resume the including file
      fclose(cur);
      cur = stack.back().fp;
      current_file = stack.back().filename;
      line_num = stack.back().lineNum;
      stack.pop_back();
    }
  }
  void Unget(int c) { if (c != EOF) pushback.push_back(c); }

  // pbrtlex.l:64-81,100-109
  void Include() {
    int c;
    do { c = Get(); } while (c == ' ' || c == '\t' || c == '\r');
    if (c != '"') ParseError("Illegal character following Include directive");
    string name;
    for (;;) {
      c = Get();
      if (c == '"') break;
      if (c == EOF || !(isalnum(c) || c == '-' || c == '_' || c == '.' || c == '/' ||
                        c == ' ' || c == '\t'))
        ParseError("Illegal character in Include file name");
      name += (char)c;
    }
    if (stack.size() > 32)
      Severe("Only 32 levels of nested Include allowed in scene files.");
    FILE *f = fopen(name.c_str(), "r");
    if (!f) Severe("Unable to open included scene file \"%s\"", name.c_str());
    Source s; s.fp = cur; s.filename = current_file; s.lineNum = line_num;
    stack.push_back(s);
    cur = f;
    current_file = name;
    line_num = 1;
  }

  // pbrtlex.l:96 NUMBER
  bool ScanNumber(int first, Token *tok) {
    string s;
    int c = first;
    if (c == '-' || c == '+') { s += (char)c; c = Get(); }
    bool digits = false;
    while (c != EOF && isdigit(c)) { s += (char)c; digits = true; c = Get(); }
    if (c == '.') {
      s += (char)c; c = Get();
      while (c != EOF && isdigit(c)) { s += (char)c; digits = true; c = Get(); }
    }
    if (!digits) { Unget(c); return false; }
    if (c == 'e' || c == 'E') {
      // exponent only if followed by [-+]?digit
      string e; e += (char)c;
      int c2 = Get();
      if (c2 == '-' || c2 == '+') { e += (char)c2; c2 = Get(); }
      if (c2 != EOF && isdigit(c2)) {
        while (c2 != EOF && isdigit(c2)) { e += (char)c2; c2 = Get(); }
        s += e;
        c = c2;
      } else {
        // not an exponent: push everything back in reverse order
        Unget(c2);
        for (size_t i = e.size(); i-- > 1;) Unget(e[i]);
        c = e[0];
      }
    }
    Unget(c);
    tok->kind = T_NUM;
    // pbrtlex.l:96: (float)atof(yytext) -- the correctly rounded double, then rounded to float.  std::from_chars
    // returns the same correctly rounded double without the locale machinery (4 M numbers in a 1 M-triangle scene);
    // it does not accept a leading '+', which atof does
#if __cplusplus >= 201703L
    {
      const char *b = s.c_str(), *e = b + s.size();
      if (b != e && *b == '+') ++b;
      double v = 0.0;
      std::from_chars_result r = std::from_chars(b, e, v);
      if (r.ec != std::errc() || r.ptr != e) v = atof(s.c_str());   // (out of range and other oddities: the C library's answer)
      tok->num = (float)v;
    }
#else
    tok->num = (float)atof(s.c_str());
#endif
    return true;
  }

  // pbrtlex.l:158-176
  void ScanString(Token *tok) {
    string s;
    for (;;) {
      int c = Get();
      if (c == EOF || c == '\n') ParseError("Unterminated string!");
      if (c == '"') break;
      if (c == '\\') {
        int e = Get();
        switch (e) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case '"': s += '"'; break;
          case '\\': s += '\\'; break;
          case '\n': line_num++; break;
          default:
            if (e != EOF && isdigit(e)) {
              int d2 = Get(), d3 = Get();
              if (d2 != EOF && isdigit(d2) && d3 != EOF && isdigit(d3)) {
                int val = (e - '0') * 100 + (d2 - '0') * 10 + (d3 - '0');
                while (val > 256) val -= 256;
                s += (char)val;
              } else {
                Unget(d3); Unget(d2);
                s += (char)e;
              }
            } else if (e != EOF)
              s += (char)e;
        }
        continue;
      }
      s += (char)c;
    }
    if (s.size() >= 1024) ParseError("String too long (1024 character limit)");
    tok->kind = T_STRING;
    tok->text = s;
  }

  Token Scan() {
    Token tok; tok.num = 0.f;
    for (;;) {
      int c = Get();
      if (c == EOF) { tok.kind = T_EOF; return tok; }
      if (c == ' ' || c == '\t' || c == '\r') continue;
      if (c == '\n') { line_num++; continue; }
      if (c == '#') {
        while ((c = Get()) != EOF && c != '\n') {}
        if (c == '\n') line_num++;
        continue;
      }
      if (c == '[') { tok.kind = T_LBRACK; return tok; }
      if (c == ']') { tok.kind = T_RBRACK; return tok; }
      if (c == '"') { ScanString(&tok); return tok; }
      if (isdigit(c) || c == '-' || c == '+' || c == '.') {
        if (ScanNumber(c, &tok)) return tok;
        ParseError("Illegal character: %c", c);
      }
      if (isalpha(c) || c == '_') {
        string id; id += (char)c;
        while ((c = Get()) != EOF && (isalnum(c) || c == '_')) id += (char)c;
        Unget(c);
        if (id == "Include") { Include(); continue; }
        tok.kind = T_KEYWORD;
        tok.text = id;
        return tok;
      }
      ParseError("Illegal character: %c", c);
    }
  }

  FILE *cur;
  vector<int> pushback;
  vector<Source> stack;
  bool havePeek;
  Token peeked;
};

// One "type name" [values] entry before typing (pbrtparse.y:262-281).
struct RawParam {
  string decl;
  vector<float> nums;
  vector<string> strs;
  bool isString;
  bool singleBareString;   // array_is_single_string (pbrtparse.y:196-199)
};

enum { PT_INT, PT_BOOL, PT_FLOAT, PT_POINT, PT_VECTOR, PT_NORMAL, PT_COLOR, PT_STRING, PT_TEXTURE };

// pbrtparse.y:543-574: the type is matched by *prefix*, in this order.
bool LookupType(const string &decl, int *type, string *name) {
  const char *strp = decl.c_str();
  while (*strp && isspace((unsigned char)*strp)) ++strp;
  if (!*strp) {
    Error("Parameter \"%s\" doesn't have a type declaration?!", decl.c_str());
    return false;
  }
  static const struct { const char *word; int type; } kinds[] = {
      {"float", PT_FLOAT},   {"integer", PT_INT},   {"bool", PT_BOOL},
      {"point", PT_POINT},   {"vector", PT_VECTOR}, {"normal", PT_NORMAL},
      {"string", PT_STRING}, {"texture", PT_TEXTURE}, {"color", PT_COLOR}};
  bool found = false;
  for (size_t i = 0; i < sizeof(kinds) / sizeof(kinds[0]); ++i) {
    size_t len = strlen(kinds[i].word);
    if (strncmp(kinds[i].word, strp, len) == 0) {
      *type = kinds[i].type;
      strp += len;
      found = true;
      break;
    }
  }
  if (!found) {
    Error("Unable to decode type for token \"%s\"", decl.c_str());
    return false;
  }
  while (*strp && isspace((unsigned char)*strp)) ++strp;
  *name = string(strp);
  return true;
}

// pbrtparse.y:470-542 (InitParamSet)
void BuildParamSet(const vector<RawParam> &raw, ParamSet *ps) {
  ps->Clear();
  for (size_t i = 0; i < raw.size(); ++i) {
    const RawParam &rp = raw[i];
    int type;
    string name;
    if (!LookupType(rp.decl, &type, &name)) {
      Warning("Type of parameter \"%s\" is unknown", rp.decl.c_str());
      continue;
    }
    if (rp.singleBareString && type != PT_TEXTURE && type != PT_STRING) {
      Warning("Bad type for %s. Changing it to a texture.", name.c_str());
      type = PT_TEXTURE;
    }
    // The reference reinterprets the raw token array as whatever the
    // declaration says; a numeric declaration fed strings (or the reverse)
    // is undefined there.  Diagnose it instead of reading garbage.
    bool wantsStrings = (type == PT_BOOL || type == PT_STRING || type == PT_TEXTURE);
    if (wantsStrings != rp.isString) {
      Error("Parameter \"%s\" has values of the wrong kind for its declared type; ignoring it.",
            rp.decl.c_str());
      continue;
    }
    int nItems = rp.isString ? (int)rp.strs.size() : (int)rp.nums.size();
    const float *f = rp.nums.empty() ? NULL : &rp.nums[0];
    switch (type) {
      case PT_INT: {
        vector<int> idata(nItems);
        for (int j = 0; j < nItems; ++j) idata[j] = int(f[j]);
        ps->AddInt(name, nItems ? &idata[0] : NULL, nItems);
        break;
      }
      case PT_BOOL: {
        // pbrtparse.y:497-510 reads element 0 for every entry; mirrored.
        bool *bdata = new bool[nItems + 1];
        for (int j = 0; j < nItems; ++j) {
          const string &s = rp.strs[0];
          if (s == "true") bdata[j] = true;
          else if (s == "false") bdata[j] = false;
          else {
            Warning("Value \"%s\" unknown for boolean parameter \"%s\"."
                    "Using \"false\".", s.c_str(), rp.decl.c_str());
            bdata[j] = false;
          }
        }
        ps->AddBool(name, bdata, nItems);
        delete[] bdata;
        break;
      }
      case PT_FLOAT:
        ps->AddFloat(name, f, nItems);
        break;
      case PT_POINT: {
        vector<Point> v;
        for (int j = 0; j + 2 < nItems; j += 3) v.push_back(Point(f[j], f[j + 1], f[j + 2]));
        ps->AddPoint(name, v.empty() ? NULL : &v[0], nItems / 3);
        break;
      }
      case PT_VECTOR: {
        vector<Vector> v;
        for (int j = 0; j + 2 < nItems; j += 3) v.push_back(Vector(f[j], f[j + 1], f[j + 2]));
        ps->AddVector(name, v.empty() ? NULL : &v[0], nItems / 3);
        break;
      }
      case PT_NORMAL: {
        vector<Normal> v;
        for (int j = 0; j + 2 < nItems; j += 3) v.push_back(Normal(f[j], f[j + 1], f[j + 2]));
        ps->AddNormal(name, v.empty() ? NULL : &v[0], nItems / 3);
        break;
      }
      case PT_COLOR: {
        vector<Spectrum> v;
        for (int j = 0; j + 2 < nItems; j += 3) {
          float c[3] = {f[j], f[j + 1], f[j + 2]};
          v.push_back(Spectrum(c));
        }
        ps->AddSpectrum(name, v.empty() ? NULL : &v[0], nItems / COLOR_SAMPLES);
        break;
      }
      case PT_STRING:
        ps->AddString(name, &rp.strs[0], nItems);
        break;
      case PT_TEXTURE:
        if (nItems == 1) ps->AddTexture(name, rp.strs[0]);
        else
          Error("Only one string allowed for \"texture\" parameter \"%s\"", name.c_str());
        break;
    }
  }
}

class Parser {
 public:
  explicit Parser(FILE *fp) : lex(fp) {}
  void Run() {
    for (;;) {
      Token t = lex.Next();
      if (t.kind == T_EOF) return;
      if (t.kind != T_KEYWORD) ParseError("syntax error");
      Statement(t.text);
    }
  }
 private:
  float Num() {
    Token t = lex.Next();
    if (t.kind != T_NUM) ParseError("syntax error");
    return t.num;
  }
  string Str() {
    Token t = lex.Next();
    if (t.kind != T_STRING) ParseError("syntax error");
    return t.text;
  }
  // array := '[' NUM+ ']' | '[' STRING+ ']' | NUM | STRING   (pbrtparse.y:176-243)
  void Array(RawParam *rp, bool allowStrings, bool bracketsRequired) {
    rp->isString = false;
    rp->singleBareString = false;
    Token t = lex.Next();
    if (t.kind == T_LBRACK) {
      Token e = lex.Next();
      if (e.kind == T_NUM) {
        while (e.kind == T_NUM) { rp->nums.push_back(e.num); e = lex.Next(); }
      } else if (e.kind == T_STRING && allowStrings) {
        rp->isString = true;
        while (e.kind == T_STRING) { rp->strs.push_back(e.text); e = lex.Next(); }
      } else
        ParseError("syntax error");
      if (e.kind != T_RBRACK) ParseError("syntax error");
    } else if (t.kind == T_NUM && !bracketsRequired) {
      rp->nums.push_back(t.num);
    } else if (t.kind == T_STRING && allowStrings && !bracketsRequired) {
      rp->isString = true;
      rp->singleBareString = true;
      rp->strs.push_back(t.text);
    } else
      ParseError("syntax error");
  }
  void ParamList(ParamSet *ps) {
    vector<RawParam> raw;
    while (lex.Peek().kind == T_STRING) {
      RawParam rp;
      rp.decl = lex.Next().text;
      Array(&rp, true, false);
      raw.push_back(rp);
    }
    BuildParamSet(raw, ps);
  }
  bool Matrix16(const char *cmd, bool bracketsRequired, float m[16]) {
    RawParam rp;
    Array(&rp, false, bracketsRequired);
    if (rp.nums.size() != 16) {
      ParseError("%s requires a(n) %d element array!", cmd, 16);
      return false;
    }
    for (int i = 0; i < 16; ++i) m[i] = rp.nums[i];
    return true;
  }
  typedef void (*NamedFn)(const string &, const ParamSet &);
  void Named(NamedFn fn) {
    string name = Str();
    ParamSet params;
    ParamList(&params);
    fn(name, params);
  }
  void Statement(const string &kw) {
    if (kw == "Accelerator") Named(pbrtAccelerator);
    else if (kw == "AreaLightSource") Named(pbrtAreaLightSource);
    else if (kw == "AttributeBegin") pbrtAttributeBegin();
    else if (kw == "AttributeEnd") pbrtAttributeEnd();
    else if (kw == "Camera") Named(pbrtCamera);
    else if (kw == "ConcatTransform") {
      float m[16];
      if (Matrix16("ConcatTransform", false, m)) pbrtConcatTransform(m);
    }
    else if (kw == "CoordinateSystem") pbrtCoordinateSystem(Str());
    else if (kw == "CoordSysTransform") pbrtCoordSysTransform(Str());
    else if (kw == "Film") Named(pbrtFilm);
    else if (kw == "Identity") pbrtIdentity();
    else if (kw == "LightSource") Named(pbrtLightSource);
    else if (kw == "LookAt") {
      float v[9];
      for (int i = 0; i < 9; ++i) v[i] = Num();
      pbrtLookAt(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
    }
    else if (kw == "Material") Named(pbrtMaterial);
    else if (kw == "ObjectBegin") pbrtObjectBegin(Str());
    else if (kw == "ObjectEnd") pbrtObjectEnd();
    else if (kw == "ObjectInstance") pbrtObjectInstance(Str());
    else if (kw == "PixelFilter") Named(pbrtPixelFilter);
    else if (kw == "ReverseOrientation") pbrtReverseOrientation();
    else if (kw == "Rotate") {
      float a = Num(), x = Num(), y = Num(), z = Num();
      pbrtRotate(a, x, y, z);
    }
    else if (kw == "Sampler") Named(pbrtSampler);
    else if (kw == "Scale") { float x = Num(), y = Num(), z = Num(); pbrtScale(x, y, z); }
    else if (kw == "SearchPath") pbrtSearchPath(Str());
    else if (kw == "Shape") Named(pbrtShape);
    else if (kw == "SurfaceIntegrator") Named(pbrtSurfaceIntegrator);
    else if (kw == "Texture") {
      string name = Str(), type = Str(), cls = Str();
      ParamSet params;
      ParamList(&params);
      pbrtTexture(name, type, cls, params);
    }
    else if (kw == "TransformBegin") pbrtTransformBegin();
    else if (kw == "TransformEnd") pbrtTransformEnd();
    else if (kw == "Transform") {
      float m[16];
      if (Matrix16("Transform", true, m)) pbrtTransform(m);
    }
    else if (kw == "Translate") { float x = Num(), y = Num(), z = Num(); pbrtTranslate(x, y, z); }
    else if (kw == "VolumeIntegrator") Named(pbrtVolumeIntegrator);
    else if (kw == "Volume") Named(pbrtVolume);
    else if (kw == "WorldBegin") pbrtWorldBegin();
    else if (kw == "WorldEnd") pbrtWorldEnd();
    else ParseError("syntax error");   // an ID token matches no rule
  }
  Lexer lex;
};

}  // namespace

// core/parser.cpp:27-51
// Data files named in a scene (image maps) are opened as given, relative to the working directory like the
// reference does; when that fails, a relative name is also tried beside the scene file being parsed.
static string gSceneDir;
string ResolveDataFile(const string &name) {
  if (name.empty() || name[0] == '/' || gSceneDir.empty()) return name;
  FILE *f = fopen(name.c_str(), "rb");
  if (f) { fclose(f); return name; }
  string alt = gSceneDir + "/" + name;
  f = fopen(alt.c_str(), "rb");
  if (f) { fclose(f); return alt; }
  return name;
}
bool ParseFile(const char *filename) {
  FILE *fp = (strcmp(filename, "-") == 0) ? stdin : fopen(filename, "r");
  if (fp != NULL) {
    string fn(filename);
    size_t slash = fn.rfind('/');
    gSceneDir = (fp != stdin && slash != string::npos) ? fn.substr(0, slash) : string("");
    current_file = filename;
    if (fp == stdin) current_file = "<standard input>";
    line_num = 1;
    {
      Parser parser(fp);
      parser.Run();
    }
    if (fp != stdin) fclose(fp);
  }
  current_file = "";
  line_num = 0;
  return (fp != NULL);
}

// Test hook (tests/test_oracle.py): the NUMBER tokens of `text`, as the tokenizer reads them (pbrtlex.l:96:
// (float)atof(yytext)).  Returns how many were written to out[0..max).
extern "C" int pbrt_host_test_scan_numbers(const char *text, float *out, int max) {
  FILE *f = fmemopen((void *)text, strlen(text), "r");
  if (!f) return -1;
  int n = 0;
  {
    Lexer lex(f);
    for (;;) {
      Token t = lex.Next();
      if (t.kind == T_EOF) break;
      if (t.kind == T_NUM && n < max) out[n++] = t.num;
    }
  }
  fclose(f);
  return n;
}
