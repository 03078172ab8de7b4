#!/bin/bash
# build_ref.sh -- compile the REFERENCE renderer (mmp/pbrt-v1) from the sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored; travels
# to the GPU box with the snapshot).  Two executables:
#
#   oracle/_ref/pbrt_ref_exact   unmodified reference sources (global MT19937)
#   oracle/_ref/pbrt_ref_keyed   reference + the per-sample RNG re-keying hooks
#                                of SURVEY.md App. H (5 inserted call lines in
#                                scene.cpp/stratified.cpp, RNG body replaced)
#
# What cannot be built as shipped, and what stands in (all under oracle/ or
# pbrt-v1_b200/host/, compiled against the reference's own headers):
#   core/exrio.cpp           needs OpenEXR      -> oracle/ref_glue.cpp (raw float dump)
#   core/pbrtlex.l, .y, parser.cpp need flex/bison -> pbrt-v1_b200/host/parser.cpp
# No reference source is copied into the repository: the two files that need
# hook lines are patched into a temporary directory that is deleted at exit.
set -euo pipefail
REF=${PBRT_REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/core" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref (if any)" >&2
  exit 0
fi
TMP=$(mktemp -d)
[ -n "${PBRT_REF_KEEP_TMP:-}" ] && echo "keeping $TMP" || trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT/plugins" "$TMP/exact" "$TMP/keyed"
# the reference's own flags (Makefile:24-25) + what g++ 13 needs to accept C++98-era code
FL="-O2 -msse2 -mfpmath=sse -I$REF -I$REF/core -I$HERE -DNDEBUG -w -fpermissive"
JOBS=${JOBS:-8}

CORE="api camera color dynload film geometry light material mc paramset primitive reflection sampling shape texture timer transform transport volume"

compile_core() {  # $1 = variant dir, $2 = extra flags
  local v=$1; shift
  for f in $CORE; do
    echo "g++ $FL $* -c $REF/core/$f.cpp -o $TMP/$v/$f.o"
  done
}

# ---- unmodified variant
{
  compile_core exact
  echo "g++ $FL -c $REF/core/util.cpp -o $TMP/exact/util.o"
} > $TMP/cmds_exact

# ---- keyed variant: util.cpp keeps its MT functions under other names
{
  compile_core keyed
  echo "g++ $FL -Dgenrand_int32=pbrt_mt_genrand_int32 -Dgenrand_real2=pbrt_mt_genrand_real2 -c $REF/core/util.cpp -o $TMP/keyed/util.o"
} > $TMP/cmds_keyed
cat $TMP/cmds_exact $TMP/cmds_keyed | xargs -P $JOBS -I{} sh -c "{}"

# ---- scene.cpp: timing/dump hooks in both variants, reseed hook is a no-op in "exact"
python3 - "$REF" "$TMP" <<'PY'
import sys, re
ref, tmp = sys.argv[1], sys.argv[2]
s = open(ref + "/core/scene.cpp").read()
decl = ('extern void pbrtRefBeginRender(); extern void pbrtRefEndRender();\n'
        'extern void pbrtRefBeforeLi(const Sampler *); extern void pbrtRefDumpSample(const Sample *, const Spectrum &, float);\n')
s = s.replace('// Scene Methods\n', decl + '// Scene Methods\n', 1)
def must(a, b, n=1):
    global s
    assert s.count(a) >= 1, a
    s = s.replace(a, b, n)
must('\twhile (sampler->GetNextSample(sample)) {', '\tpbrtRefBeginRender();\n\twhile (sampler->GetNextSample(sample)) {')
must('\t\tif (rayWeight > 0.f)', '\t\tpbrtRefBeforeLi(sampler);\n\t\tif (rayWeight > 0.f)')
must('\t\t// Add sample contribution to image', '\t\tpbrtRefDumpSample(sample, Ls, alpha);\n\t\t// Add sample contribution to image')
must('\t// Clean up after rendering and store final image', '\tpbrtRefEndRender();\n\t// Clean up after rendering and store final image')
open(tmp + "/scene_hooked.cpp", "w").write(s)
t = open(ref + "/samplers/stratified.cpp").read()
t = t.replace('// StratifiedSampler Declarations', 'extern void pbrtReseed(int, int, int, int);\n// StratifiedSampler Declarations', 1)
lines = t.split('\n')
out = []
n_cam = n_int = 0
for ln in lines:
    if ln.strip() == '// Generate stratified camera samples for (_xPos_,_yPos_)':
        out.append(ln[:len(ln) - len(ln.lstrip())] + 'pbrtReseed(xPos, yPos, 0xFFFF, 0);'); n_cam += 1
    if ln.strip() == '// Generate stratified samples for integrators':
        out.append(ln[:len(ln) - len(ln.lstrip())] + 'pbrtReseed(xPos, yPos, samplePos, 0);'); n_int += 1
    out.append(ln)
assert n_cam == 2 and n_int == 1, (n_cam, n_int)
t = '\n'.join(out)
open(tmp + "/stratified_keyed.cpp", "w").write(t)
# lowdiscrepancy.cpp: one re-key per pixel, before its (0,2)-sequence tables are generated
l = open(ref + "/samplers/lowdiscrepancy.cpp").read()
l = l.replace('// LDSampler Declarations', 'extern void pbrtReseed(int, int, int, int);\n// LDSampler Declarations', 1)
mark = '\t\t// Generate low-discrepancy samples for pixel\n'
assert l.count(mark) == 1
l = l.replace(mark, '\t\tpbrtReseed(xPos, yPos, 0xFFFF, 0);\n' + mark)
open(tmp + "/lowdiscrepancy_keyed.cpp", "w").write(l)
# bestcandidate.cpp: one re-key per table placement (sample shifts + LDShuffleScrambled tables) and one
# per sample (its stratified patterns; also identifies the sample for the Li re-key)
b = open(ref + "/samplers/bestcandidate.cpp").read()
b = b.replace('// BestCandidateSampler Declarations', 'extern void pbrtReseed(int, int, int, int);\nextern void pbrtRefBestCandidateSample(int, int, int);\n// BestCandidateSampler Declarations', 1)
m1 = '\t\t// Update sample shifts\n'
m2 = '\t// Compute integrator samples for best-candidate sample\n'
assert b.count(m1) == 1 and b.count(m2) == 1
b = b.replace(m1, '\t\tpbrtReseed(Floor2Int(xTableCorner), Floor2Int(yTableCorner), 0, 3);\n' + m1)
b = b.replace(m2, '\tpbrtRefBestCandidateSample(Floor2Int(xTableCorner), Floor2Int(yTableCorner), tableOffset);\n' + m2)
open(tmp + "/bestcandidate_keyed.cpp", "w").write(b)
# random.cpp: same two re-key points as the stratified sampler (pixel pattern: constructor and
# GetNextSample; integrator samples of each image sample)
r = open(ref + "/samplers/random.cpp").read()
r = r.replace('class RandomSampler : public Sampler {', 'extern void pbrtReseed(int, int, int, int);\nclass RandomSampler : public Sampler {', 1)
lines = r.split('\n')
out = []
n_pix = n_int = 0
for i, ln in enumerate(lines):
    if ln.strip() == 'for (int i = 0;' and lines[i + 1].strip() == 'i < 5 * xPixelSamples * yPixelSamples;':
        out.append(ln[:len(ln) - len(ln.lstrip())] + 'pbrtReseed(xPos, yPos, 0xFFFF, 0);'); n_pix += 1
    if ln.strip() == '// Generate stratified samples for integrators':
        out.append(ln[:len(ln) - len(ln.lstrip())] + 'pbrtReseed(xPos, yPos, samplePos, 0);'); n_int += 1
    out.append(ln)
assert n_pix == 2 and n_int == 1, (n_pix, n_int)
open(tmp + "/random_keyed.cpp", "w").write('\n'.join(out))
# kdtree.cpp: count calls reaching the kd walk (the metric's "rays": Scene::Intersect + IntersectP)
k = open(ref + "/accelerators/kdtree.cpp").read()
k = k.replace('// KdAccelNode Declarations', 'extern unsigned long long gRefIntersectCalls, gRefIntersectPCalls;\n// KdAccelNode Declarations', 1)
a = 'bool KdTreeAccel::Intersect(const Ray &ray,\n\t\tIntersection *isect) const {\n'
b = 'bool KdTreeAccel::IntersectP(const Ray &ray) const {\n'
assert k.count(a) == 1 and k.count(b) == 1
k = k.replace(a, a + '\t++gRefIntersectCalls;\n').replace(b, b + '\t++gRefIntersectPCalls;\n')
assert 'extern unsigned long long gRefIntersectCalls' in k
# ... and hand the finished tree to the glue (exact node / reference counts on stderr; with
# PBRT_REF_DUMP_KD=<file> the whole tree: tests compare the host build's arrays with it word for word)
k = k.replace('// KdAccelNode Declarations', 'extern void pbrtRefKdBuilt(const void *nodes, int nNodes, const void *mailboxBase, int mailboxStride, int nPrims, const float *bounds);\n// KdAccelNode Declarations', 1)
m = '\t// Free working memory for kd-tree construction\n'
assert k.count(m) == 1
k = k.replace(m, '\t{ float b6[6] = { bounds.pMin.x, bounds.pMin.y, bounds.pMin.z, bounds.pMax.x, bounds.pMax.y, bounds.pMax.z };\n'
                 '\t  pbrtRefKdBuilt(nodes, nextFreeNode, mailboxPrims, (int)sizeof(MailboxPrim), (int)nMailboxes, b6); }\n' + m)
open(tmp + "/kdtree_counted.cpp", "w").write(k)
PY
g++ $FL -c $TMP/scene_hooked.cpp -o $TMP/exact/scene.o
cp $TMP/exact/scene.o $TMP/keyed/scene.o
g++ $FL -c $HERE/ref_glue.cpp -o $TMP/exact/ref_glue.o
g++ $FL -DPBRT_REF_KEYED -c $HERE/ref_glue.cpp -o $TMP/keyed/ref_glue.o
cp $HERE/../pbrt-v1_b200/host/parser.cpp $TMP/ref_parser.cpp   # own source; moved so "paramset.h" resolves to the reference's
g++ $FL -c $TMP/ref_parser.cpp -o $TMP/parser.o
g++ $FL -c $REF/renderer/pbrt.cpp -o $TMP/main.o

# ---- plugins (every directory the reference Makefile lists; sampledata.cpp is an include file)
PLUGDIRS="accelerators cameras film filters integrators lights materials samplers shapes textures tonemaps volumes"
: > $TMP/cmds_plug
for d in $PLUGDIRS; do
  for src in $REF/$d/*.cpp; do
    b=$(basename $src .cpp)
    [ "$b" = "sampledata" ] && continue
    echo "g++ $FL -fPIC -shared $src -o $OUT/plugins/$b.so" >> $TMP/cmds_plug
  done
done
xargs -P $JOBS -I{} sh -c "{}" < $TMP/cmds_plug
g++ $FL -I$REF/accelerators -fPIC -shared $TMP/kdtree_counted.cpp -o $OUT/plugins/kdtree.so
mkdir -p "$OUT/plugins_keyed"
for f in $OUT/plugins/*.so; do ln -sf ../plugins/$(basename $f) $OUT/plugins_keyed/$(basename $f); done
rm -f $OUT/plugins_keyed/stratified.so
g++ $FL -fPIC -shared $TMP/stratified_keyed.cpp -o $OUT/plugins_keyed/stratified.so
rm -f $OUT/plugins_keyed/lowdiscrepancy.so
g++ $FL -fPIC -shared $TMP/lowdiscrepancy_keyed.cpp -o $OUT/plugins_keyed/lowdiscrepancy.so
rm -f $OUT/plugins_keyed/random.so
g++ $FL -fPIC -shared $TMP/random_keyed.cpp -o $OUT/plugins_keyed/random.so
rm -f $OUT/plugins_keyed/bestcandidate.so
g++ $FL -I$REF -fPIC -shared $TMP/bestcandidate_keyed.cpp -o $OUT/plugins_keyed/bestcandidate.so

g++ -O2 -msse2 -mfpmath=sse -rdynamic -o $OUT/pbrt_ref_exact $TMP/main.o $TMP/parser.o $TMP/exact/*.o -ldl -lm
g++ -O2 -msse2 -mfpmath=sse -rdynamic -o $OUT/pbrt_ref_keyed $TMP/main.o $TMP/parser.o $TMP/keyed/*.o -ldl -lm
echo "built: $OUT/pbrt_ref_exact $OUT/pbrt_ref_keyed ($(ls $OUT/plugins | wc -l) plugins)"
