// Native dataset ingest: the text of a data.json (one sample object or an array of them) -> the
// block-diagonal batch arrays the device path starts from (src / dst / seq per adjacency, features,
// labels, per-sample entity offsets).
//
// It replaces, on the host side, the per-edge Python loops of the reference's generator
// (code/utils/generator_std_to_framework.py:32-50 make_indices, :134-190 adjacency loop, :102-107
// features) and the concatenation of per-sample tensors: SURVEY.md section 8f rank 1 ("once the kernels
// are fast the Python per-edge append loop is the end-to-end bottleneck").  Results are identical,
// array for array, to ignnition_b200/generator.py + batching.assemble (tests/test_host.py).  Host C++
// only: no CUDA in this file.

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/ignnition_b200.h"

void ign_set_error(const char* fmt, ...);

namespace {

struct Err {
  std::string msg;
};

// ------------------------------------------------------------------------------------------
// a JSON scanner over a byte range (no DOM: the schema is walked directly)
// ------------------------------------------------------------------------------------------
struct Scan {
  const char* p;
  const char* end;

  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  char peek() {
    ws();
    if (p >= end) throw Err{"IGNNITION: unexpected end of the dataset text"};
    return *p;
  }
  void expect(char c) {
    if (peek() != c) throw Err{std::string("IGNNITION: malformed data.json: expected '") + c + "' near \"" +
                               std::string(p, (size_t)std::min<ptrdiff_t>(end - p, 24)) + "\""};
    ++p;
  }
  bool accept(char c) {
    if (peek() == c) { ++p; return true; }
    return false;
  }
  // string without escapes -> view into the text; with escapes -> unescaped copy kept in `arena`
  std::string_view str(std::deque<std::string>& arena) {
    expect('"');
    const char* s = p;
    bool esc = false;
    while (p < end && *p != '"') {
      if (*p == '\\') { esc = true; ++p; }
      ++p;
    }
    if (p >= end) throw Err{"IGNNITION: malformed data.json: unterminated string"};
    std::string_view v(s, (size_t)(p - s));
    ++p;
    if (!esc) return v;
    std::string out;
    for (size_t i = 0; i < v.size(); ++i) {
      if (v[i] != '\\') { out.push_back(v[i]); continue; }
      ++i;
      switch (v[i]) {
        case 'n': out.push_back('\n'); break;
        case 't': out.push_back('\t'); break;
        case 'r': out.push_back('\r'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'u': {            // \uXXXX -> UTF-8 (basic multilingual plane; json.dumps escapes non-ASCII this way)
          unsigned cp = 0;
          for (int k = 1; k <= 4 && i + k < v.size(); ++k) {
            const char c = v[i + k];
            cp = cp * 16 + (c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
          }
          i += 4;
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 63))); }
          else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 63))); out.push_back((char)(0x80 | (cp & 63))); }
          break;
        }
        default: out.push_back(v[i]);
      }
    }
    arena.push_back(std::move(out));
    return arena.back();
  }
  double number() {
    ws();
    double d = 0.0;
    auto r = std::from_chars(p, end, d);
    if (r.ec != std::errc()) {
      // json.dumps writes NaN / Infinity / -Infinity for non-finite floats
      if (end - p >= 3 && !strncmp(p, "NaN", 3)) { p += 3; return std::nan(""); }
      if (end - p >= 8 && !strncmp(p, "Infinity", 8)) { p += 8; return INFINITY; }
      if (end - p >= 9 && !strncmp(p, "-Infinity", 9)) { p += 9; return -INFINITY; }
      throw Err{std::string("IGNNITION: malformed data.json: a number was expected near \"") +
                std::string(p, (size_t)std::min<ptrdiff_t>(end - p, 24)) + "\""};
    }
    p = r.ptr;
    return d;
  }
  void skip() {                 // any value
    const char c = peek();
    if (c == '"') {
      ++p;
      while (p < end && *p != '"') { if (*p == '\\') ++p; ++p; }
      ++p;
    } else if (c == '{' || c == '[') {
      int depth = 0;
      do {
        const char d = *p;
        if (d == '"') {
          ++p;
          while (p < end && *p != '"') { if (*p == '\\') ++p; ++p; }
        } else if (d == '{' || d == '[') ++depth;
        else if (d == '}' || d == ']') --depth;
        ++p;
      } while (depth > 0 && p < end);
      if (depth) throw Err{"IGNNITION: malformed data.json: unbalanced brackets"};
    } else {
      while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') ++p;
    }
  }
};

// ------------------------------------------------------------------------------------------
// name -> (entity type, index): open addressing over views into the text
// ------------------------------------------------------------------------------------------
struct NameMap {
  struct Slot { std::string_view key; int32_t type; int32_t index; };
  std::vector<Slot> slots;
  size_t mask = 0, used = 0;
  static uint64_t hash(std::string_view s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h ^ (h >> 29);
  }
  void reset(size_t expect) {
    size_t cap = 64;
    while (cap < 2 * expect) cap <<= 1;
    if (slots.size() < cap) slots.resize(cap);
    mask = cap - 1;
    for (size_t i = 0; i < cap; ++i) slots[i].type = -1;
    used = 0;
  }
  void grow() {
    std::vector<Slot> old(slots.begin(), slots.begin() + (mask + 1));
    reset(old.size());
    for (auto& s : old) if (s.type >= 0) put(s.key, s.type, s.index);
  }
  void put(std::string_view k, int32_t type, int32_t index) {
    if (2 * (used + 1) > mask + 1) grow();
    size_t i = hash(k) & mask;
    while (slots[i].type >= 0 && slots[i].key != k) i = (i + 1) & mask;
    if (slots[i].type < 0) ++used;
    slots[i] = Slot{k, type, index};            // a repeated name keeps the last value, like a Python dict
  }
  const Slot* get(std::string_view k) const {
    size_t i = hash(k) & mask;
    while (slots[i].type >= 0) {
      if (slots[i].key == k) return &slots[i];
      i = (i + 1) & mask;
    }
    return nullptr;
  }
};

struct Adj {
  std::string name;
  int src, dst;
  bool params;
  std::vector<int32_t> s, d, q;
  std::vector<float> p;
  int32_t p_width = 0;
};
struct Feat {
  std::string name;
  int entity;
  std::vector<float> v;
};
// a multi-source ordered / interleave message passing: which (source, column) of the reference's
// concatenated padded tensor sits at sequence position p (batching.position_table)
struct Seq {
  std::vector<int> adjs;
  bool interleave;
  std::string pattern_key;           // sample key of the interleave pattern (list of entity type names)
  std::vector<int32_t> off{0}, src, col;
  std::vector<int> pattern;          // this sample's pattern as entity type indices
  bool have_pattern = false;
};
constexpr int32_t BIG_COL = 1 << 30;   // "no such column" (batching.BIG_COL)

}  // namespace

struct ign_ingest {
  std::vector<std::string> entity_names;
  std::vector<Feat> feats;
  std::vector<Adj> adjs;
  std::vector<Seq> seqs;
  std::string label;
  bool has_label = false;
  std::vector<float> labels;
  std::vector<std::vector<int64_t>> offsets;        // per entity type: [n_samples + 1]
  int64_t n_samples = 0;
  // per-sample scratch
  NameMap names;
  std::deque<std::string> arena;                    // unescaped names (stable addresses)
  size_t last_entities = 1024;                       // sizes the name table of the next sample

  void reset() {
    for (auto& f : feats) f.v.clear();
    for (auto& a : adjs) { a.s.clear(); a.d.clear(); a.q.clear(); a.p.clear(); a.p_width = 0; }
    labels.clear();
    for (auto& q : seqs) { q.off.assign(1, 0); q.src.clear(); q.col.clear(); }
    for (auto& o : offsets) o.assign(1, 0);
    n_samples = 0;
  }

  int entity_index(std::string_view type) const {
    for (size_t i = 0; i < entity_names.size(); ++i)
      if (entity_names[i] == type) return (int)i;
    return -1;
  }

  // flattens a number or (nested) list of numbers
  static void numbers(Scan& sc, std::vector<float>& out, bool truncate) {
    if (sc.peek() == '[') {
      ++sc.p;
      if (sc.accept(']')) return;
      do numbers(sc, out, truncate); while (sc.accept(','));
      sc.expect(']');
    } else {
      const double d = sc.number();
      out.push_back((float)(truncate ? std::trunc(d) : d));
    }
  }

  // ---- make_indices (generator_std_to_framework.py:32-50): per-type counters in order of appearance
  std::vector<int64_t> count;
  std::vector<std::string_view> other_types;                    // types the model does not use
  std::vector<int64_t> other_count;
  std::string type_name(int t) const {
    return t < (int)entity_names.size() ? entity_names[t] : std::string(other_types[t - entity_names.size()]);
  }
  void parse_entities(Scan& e) {
    count.assign(entity_names.size(), 0);
    other_types.clear();
    other_count.clear();
    e.expect('{');
    names.reset(last_entities);
    if (!e.accept('}')) {
      do {
        std::string_view node = e.str(arena);
        e.expect(':');
        std::string_view type = e.str(arena);
        int t = entity_index(type);
        if (t >= 0) {
          names.put(node, t, (int32_t)count[t]++);
        } else {
          size_t j = 0;
          while (j < other_types.size() && other_types[j] != type) ++j;
          if (j == other_types.size()) { other_types.push_back(type); other_count.push_back(0); }
          names.put(node, (int32_t)(entity_names.size() + j), (int32_t)other_count[j]++);
        }
      } while (e.accept(','));
      e.expect('}');
    }
    last_entities = names.used > 64 ? names.used : 64;
  }

  // ---- one adjacency (:134-190), grouped by destination in the order of the text
  void parse_adjacency(Adj& a, Scan& s) {
    const int64_t so = offsets[a.src].back(), dofs = offsets[a.dst].back();
    s.expect('{');
    if (s.accept('}')) return;
    do {
      std::string_view dname = s.str(arena);
      const NameMap::Slot* d = names.get(dname);
      if (!d) throw Err{"IGNNITION: \"" + std::string(dname) + "\" of the adjecency list \"" + a.name + "\" is not in the entities of its sample"};
      if (d->type != a.dst)
        throw Err{"IGNNITION: The adjecency list \"" + a.name + "\" was expected to be from " + entity_names[a.src] + " to " +
                  entity_names[a.dst] + ".\n However, \"" + std::string(dname) + "\" was found which is of type \"" +
                  type_name(d->type) + "\" instead of " + entity_names[a.dst]};
      s.expect(':');
      s.expect('[');
      int32_t pos = 0;
      if (!s.accept(']')) {
        do {
          std::string_view sname;
          if (s.peek() == '[') {                      // [name, parameters]
            ++s.p;
            sname = s.str(arena);
            if (s.accept(',')) {
              if (a.params) {
                const size_t before = a.p.size();
                numbers(s, a.p, true);                // declared int64 then cast: truncation (quirk 11)
                const int32_t w = (int32_t)(a.p.size() - before);
                if (a.p_width == 0) a.p_width = w;
                else if (w != a.p_width) throw Err{"IGNNITION: edge parameters of \"" + a.name + "\" change width inside the dataset"};
              } else {
                s.skip();
              }
              while (s.accept(',')) s.skip();
            }
            s.expect(']');
            const NameMap::Slot* e = names.get(sname);
            if (!e) throw Err{"IGNNITION: \"" + std::string(sname) + "\" of the adjecency list \"" + a.name + "\" is not in the entities of its sample"};
            a.s.push_back((int32_t)(so + e->index));
          } else {
            sname = s.str(arena);
            const NameMap::Slot* e = names.get(sname);
            if (!e) throw Err{"IGNNITION: \"" + std::string(sname) + "\" of the adjecency list \"" + a.name + "\" is not in the entities of its sample"};
            if (e->type != a.src)
              throw Err{"IGNNITION: The adjecency list \"" + a.name + "\" was expected to be from \"" + entity_names[a.src] +
                        "\" to \"" + entity_names[a.dst] + ".\n However, \"" + std::string(sname) + "\" was found which is of type \"" +
                        type_name(e->type) + "\" instead of \"" + entity_names[a.src]};
            a.s.push_back((int32_t)(so + e->index));
          }
          a.d.push_back((int32_t)(dofs + d->index));
          a.q.push_back(pos++);
        } while (s.accept(','));
        s.expect(']');
      }
    } while (s.accept(','));
    s.expect('}');
  }

  // One sample object, in ONE pass over its text when "entities" comes before the adjacency lists (the
  // adjacency lists met earlier are revisited once the name table exists).  A repeated key replaces the
  // earlier value, like json.loads.
  void sample(Scan& sc) {
    sc.expect('{');
    arena.clear();
    const size_t nf = feats.size(), na = adjs.size();
    std::vector<size_t> f0(nf), a0(na), ap0(na);
    for (size_t i = 0; i < nf; ++i) f0[i] = feats[i].v.size();
    for (size_t i = 0; i < na; ++i) { a0[i] = adjs[i].s.size(); ap0[i] = adjs[i].p.size(); }
    const size_t l0 = labels.size();
    for (auto& q : seqs) q.have_pattern = false;
    std::vector<char> f_seen(nf, 0), a_done(na, 0);
    std::vector<const char*> a_at(na, nullptr);
    bool have_entities = false, have_label = false;
    auto rollback_adj = [&](size_t i) {
      adjs[i].s.resize(a0[i]); adjs[i].d.resize(a0[i]); adjs[i].q.resize(a0[i]); adjs[i].p.resize(ap0[i]);
    };
    if (!sc.accept('}')) {
      do {
        std::string_view k = sc.str(arena);
        sc.expect(':');
        sc.ws();
        bool used = false;
        if (k == "entities") {
          parse_entities(sc);
          have_entities = true;
          for (size_t i = 0; i < na; ++i)            // adjacency lists parsed against an earlier "entities": redo
            if (a_done[i]) { rollback_adj(i); a_done[i] = 0; }
          used = true;
        }
        for (size_t i = 0; i < nf && !used; ++i)
          if (k == feats[i].name) {
            feats[i].v.resize(f0[i]);
            numbers(sc, feats[i].v, false);
            f_seen[i] = 1;
            used = true;
          }
        if (!used && has_label && k == label) {
          labels.resize(l0);
          numbers(sc, labels, false);
          have_label = true;
          used = true;
        }
        for (size_t i = 0; i < na && !used; ++i)
          if (k == adjs[i].name) {
            a_at[i] = sc.p;
            if (a_done[i]) { rollback_adj(i); a_done[i] = 0; }
            if (have_entities) {
              parse_adjacency(adjs[i], sc);
              a_done[i] = 1;
            } else {
              sc.skip();
            }
            used = true;
          }
        for (size_t i = 0; i < seqs.size() && !used; ++i)
          if (seqs[i].interleave && k == seqs[i].pattern_key) {
            const char* at = sc.p;
            for (size_t j = i; j < seqs.size(); ++j) {        // several message passings may share one definition
              if (!seqs[j].interleave || seqs[j].pattern_key != k) continue;
              Scan ps{at, sc.end};
              seqs[j].pattern.clear();
              ps.expect('[');
              if (!ps.accept(']')) {
                do {
                  std::string_view ent = ps.str(arena);
                  const int t = entity_index(ent);
                  if (t < 0) throw Err{"IGNNITION: the interleave definition \"" + seqs[j].pattern_key + "\" names the unknown entity \"" + std::string(ent) + "\""};
                  seqs[j].pattern.push_back(t);
                } while (ps.accept(','));
                ps.expect(']');
              }
              seqs[j].have_pattern = true;
              sc.p = ps.p;
            }
            used = true;
          }
        if (!used) sc.skip();
      } while (sc.accept(','));
      sc.expect('}');
    }
    const char* after = sc.p;
    if (!have_entities) throw Err{"IGNNITION: a sample without an \"entities\" dictionary was found"};
    for (size_t i = 0; i < nf; ++i)
      if (!f_seen[i]) throw Err{"IGNNITION: A list for feature named \"" + feats[i].name + "\" was not found although being expected."};
    if (has_label && !have_label)
      throw Err{"IGNNITION: A list for the output named \"" + label + "\" was not found although being expected."};
    for (size_t i = 0; i < na; ++i) {
      if (a_done[i]) continue;
      if (!a_at[i]) throw Err{"IGNNITION: A list for the adjecency vector named \"" + adjs[i].name + "\" was not found although being expected."};
      Scan s{a_at[i], sc.end};
      parse_adjacency(adjs[i], s);
    }
    // ---- position tables of the multi-source sequences (batching.position_table; interleave:
    // generator_std_to_framework.py:193-219 + auxilary_classes.py:421-440)
    for (auto& q : seqs) {
      std::vector<int32_t> maxlen(q.adjs.size(), 0);
      int64_t total = 0;
      for (size_t k = 0; k < q.adjs.size(); ++k) {
        const Adj& a = adjs[(size_t)q.adjs[k]];
        int32_t m = 0;
        for (size_t j = a0[(size_t)q.adjs[k]]; j < a.q.size(); ++j) m = a.q[j] + 1 > m ? a.q[j] + 1 : m;
        maxlen[k] = m;
        total += m;
      }
      const size_t base = q.src.size();
      q.src.resize(base + (size_t)total, 0);
      q.col.resize(base + (size_t)total, q.interleave ? BIG_COL : 0);
      if (!q.interleave) {
        size_t p = base;
        for (size_t k = 0; k < q.adjs.size(); ++k)
          for (int32_t c = 0; c < maxlen[k]; ++c, ++p) { q.src[p] = (int32_t)k; q.col[p] = c; }
      } else {
        if (!q.have_pattern) throw Err{"IGNNITION: the interleave definition \"" + q.pattern_key + "\" was not found although being expected."};
        // ids of the pattern's entities in order of first appearance; n_total = sum of their longest lists
        std::vector<int> ids;                     // entity type of id i
        std::vector<int> numeric;
        for (int t : q.pattern) {
          size_t i = 0;
          while (i < ids.size() && ids[i] != t) ++i;
          if (i == ids.size()) ids.push_back(t);
          numeric.push_back((int)i);
        }
        std::vector<int32_t> ent_max(ids.size(), -1);
        for (size_t i = 0; i < ids.size(); ++i)
          for (size_t k = 0; k < q.adjs.size(); ++k)
            if (adjs[(size_t)q.adjs[k]].src == ids[i]) ent_max[i] = maxlen[k];
        int64_t n_total = 0;
        for (size_t i = 0; i < ids.size(); ++i) {
          if (ent_max[i] < 0) throw Err{"IGNNITION: the interleave definition \"" + q.pattern_key + "\" names \"" + entity_names[(size_t)ids[i]] + "\", which sends no message here"};
          n_total += ent_max[i];
        }
        // positions of entity i in the pattern tiled to n_total entries, then concatenated in source order
        std::vector<std::vector<int32_t>> where(ids.size());
        if (!numeric.empty())
          for (int64_t p = 0; p < n_total; ++p) where[(size_t)numeric[(size_t)(p % (int64_t)numeric.size())]].push_back((int32_t)p);
        size_t j = 0;
        bool ok = true;
        for (size_t k = 0; k < q.adjs.size() && ok; ++k) {
          size_t i = 0;
          while (i < ids.size() && ids[i] != adjs[(size_t)q.adjs[k]].src) ++i;
          if (i == ids.size()) throw Err{"IGNNITION: the interleave definition \"" + q.pattern_key + "\" does not name \"" + entity_names[(size_t)adjs[(size_t)q.adjs[k]].src] + "\""};
          // column c of source k's block (k-th block of the concatenation) moves to position where[i][c']
          for (int32_t w : where[i]) {
            if ((int64_t)j >= total) { ok = false; break; }
            // j-th entry of the concatenated padded tensor: (source, column)
            size_t kk = 0; int64_t jj = (int64_t)j;
            while (kk < maxlen.size() && jj >= maxlen[kk]) { jj -= maxlen[kk]; ++kk; }
            if (w >= total) { ok = false; break; }
            q.src[base + (size_t)w] = (int32_t)kk;
            q.col[base + (size_t)w] = (int32_t)jj;
            ++j;
          }
        }
        if (!ok || (int64_t)j != total)
          throw Err{"IGNNITION: interleave indices (" + std::to_string(n_total) + ") do not match the padded length (" + std::to_string(total) + ")"};
      }
      q.off.push_back(q.off.back() + (int32_t)total);
    }
    for (size_t t = 0; t < entity_names.size(); ++t) {
      const int64_t total = offsets[t].back() + count[t];
      if (total >= ((int64_t)1 << 31)) throw Err{"IGNNITION: more than 2^31 entities of type " + entity_names[t] + " in one batch"};
      offsets[t].push_back(total);
    }
    ++n_samples;
    sc.p = after;
  }

};

extern "C" ign_ingest_t* ign_ingest_create(int n_entities, const char* const* entity_names, int n_features,
                                           const char* const* feature_names, const int32_t* feature_entity,
                                           int n_adj, const char* const* adj_names, const int32_t* adj_src,
                                           const int32_t* adj_dst, const int32_t* adj_params, const char* label_name) {
  if (n_entities <= 0 || n_features < 0 || n_adj < 0 || !entity_names) {
    ign_set_error("IGNNITION: ingest_create: bad argument");
    return nullptr;
  }
  ign_ingest* g = new ign_ingest();
  for (int i = 0; i < n_entities; ++i) g->entity_names.emplace_back(entity_names[i]);
  for (int i = 0; i < n_features; ++i) {
    if (feature_entity[i] < 0 || feature_entity[i] >= n_entities) {
      ign_set_error("IGNNITION: ingest_create: feature %s belongs to no entity", feature_names[i]);
      delete g;
      return nullptr;
    }
    g->feats.push_back(Feat{feature_names[i], feature_entity[i], {}});
  }
  for (int i = 0; i < n_adj; ++i) {
    if (adj_src[i] < 0 || adj_src[i] >= n_entities || adj_dst[i] < 0 || adj_dst[i] >= n_entities) {
      ign_set_error("IGNNITION: ingest_create: adjacency %s joins unknown entities", adj_names[i]);
      delete g;
      return nullptr;
    }
    Adj a;
    a.name = adj_names[i]; a.src = adj_src[i]; a.dst = adj_dst[i]; a.params = adj_params && adj_params[i] != 0;
    g->adjs.push_back(std::move(a));
  }
  if (label_name) { g->label = label_name; g->has_label = true; }
  g->offsets.assign((size_t)n_entities, std::vector<int64_t>(1, 0));
  return g;
}

extern "C" int ign_ingest_add_sequence(ign_ingest_t* g, int n_adj, const int32_t* adj_indices, int interleave,
                                       const char* pattern_key) {
  if (!g || n_adj <= 0 || !adj_indices || (interleave && !pattern_key)) {
    ign_set_error("IGNNITION: ingest_add_sequence: bad argument");
    return IGN_ERR_INVALID;
  }
  Seq q;
  for (int i = 0; i < n_adj; ++i) {
    if (adj_indices[i] < 0 || adj_indices[i] >= (int)g->adjs.size()) {
      ign_set_error("IGNNITION: ingest_add_sequence: unknown adjacency %d", adj_indices[i]);
      return IGN_ERR_INVALID;
    }
    q.adjs.push_back(adj_indices[i]);
  }
  q.interleave = interleave != 0;
  if (pattern_key) q.pattern_key = pattern_key;
  g->seqs.push_back(std::move(q));
  return (int)g->seqs.size() - 1;
}

extern "C" int64_t ign_ingest_sequence(const ign_ingest_t* g, int seq, const int32_t** pos_off, const int32_t** pos_src,
                                       const int32_t** pos_col) {
  if (!g || seq < 0 || seq >= (int)g->seqs.size()) return IGN_ERR_INVALID;
  const Seq& q = g->seqs[(size_t)seq];
  if (pos_off) *pos_off = q.off.data();
  if (pos_src) *pos_src = q.src.data();
  if (pos_col) *pos_col = q.col.data();
  return (int64_t)q.src.size();
}

extern "C" void ign_ingest_destroy(ign_ingest_t* g) { delete g; }
extern "C" void ign_ingest_reset(ign_ingest_t* g) { if (g) g->reset(); }

extern "C" int64_t ign_ingest_parse(ign_ingest_t* g, const char* json, size_t len, int64_t max_samples) {
  if (!g || (!json && len)) {
    ign_set_error("IGNNITION: ingest_parse: null pointer");
    return IGN_ERR_INVALID;
  }
  Scan sc{json, json + len};
  int64_t done = 0;
  try {
    if (sc.peek() != '[') {
      g->sample(sc);
      return 1;
    }
    ++sc.p;
    if (!sc.accept(']')) {
      do {
        if (max_samples >= 0 && done >= max_samples) break;
        g->sample(sc);
        ++done;
      } while (sc.accept(','));
    }
    return done;
  } catch (const Err& e) {
    ign_set_error("%s", e.msg.c_str());
    return IGN_ERR_INVALID;
  } catch (const std::exception& e) {
    ign_set_error("IGNNITION: ingest_parse: %s", e.what());
    return IGN_ERR_INVALID;
  }
}

extern "C" int64_t ign_ingest_n_samples(const ign_ingest_t* g) { return g ? g->n_samples : 0; }

extern "C" const int64_t* ign_ingest_offsets(const ign_ingest_t* g, int entity) {
  if (!g || entity < 0 || entity >= (int)g->offsets.size()) return nullptr;
  return g->offsets[entity].data();
}

extern "C" int64_t ign_ingest_feature(const ign_ingest_t* g, int feature, const float** data) {
  if (!g || feature < 0 || feature >= (int)g->feats.size()) return IGN_ERR_INVALID;
  if (data) *data = g->feats[feature].v.data();
  return (int64_t)g->feats[feature].v.size();
}

extern "C" int64_t ign_ingest_adjacency(const ign_ingest_t* g, int adj, const int32_t** src, const int32_t** dst,
                                        const int32_t** seq, const float** params, int32_t* params_width) {
  if (!g || adj < 0 || adj >= (int)g->adjs.size()) return IGN_ERR_INVALID;
  const Adj& a = g->adjs[adj];
  if (src) *src = a.s.data();
  if (dst) *dst = a.d.data();
  if (seq) *seq = a.q.data();
  if (params) *params = a.p.data();
  if (params_width) *params_width = a.p_width;
  return (int64_t)a.s.size();
}

extern "C" int64_t ign_ingest_labels(const ign_ingest_t* g, const float** data) {
  if (!g) return IGN_ERR_INVALID;
  if (data) *data = g->labels.data();
  return (int64_t)g->labels.size();
}
