// loco_proto.cpp -- hand-written proto2 codec for OptCAD.proto.PrecomputedParametricShape.
//
// Field numbers / names follow LCFunction/PrecomputedParametricShape.proto:4-141; field usage and
// load order follow LCProtoConverter.cpp:141-168 (writer) and :356-403 (reader).  Both the binary
// wire format and the protobuf text format go through one small DOM driven by a schema table.
#include "loco_proto.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <unordered_map>

namespace loco {
namespace {

// ---- schema ------------------------------------------------------------------------------------
enum MsgType { M_SHAPE, M_SAMPLE, M_SHAPEINFO, M_FTEST, M_TETMESH, M_VEC3, M_BASIS, M_BSPLINE, M_CELL, M_LEAF, M_HSAMPLE, M_INTER, M_COUNT };
enum Kind { K_INT32, K_DOUBLE, K_MSG };
struct FieldDef { int num; const char *name; Kind kind; MsgType sub; bool repeated; };

const std::vector<FieldDef> &schema(MsgType t)
{
	static const std::vector<FieldDef> S[M_COUNT] = {
		/* M_SHAPE     proto:4-10   */ {{1, "root", K_MSG, M_CELL, false}, {2, "samples", K_MSG, M_SAMPLE, true}, {3, "bootstrapNSamples", K_INT32, M_COUNT, false}, {4, "borderCells", K_MSG, M_CELL, true}, {5, "range", K_DOUBLE, M_COUNT, true}},
		/* M_SAMPLE    proto:12-17  */ {{1, "id", K_INT32, M_COUNT, false}, {2, "center", K_DOUBLE, M_COUNT, true}, {3, "shapeInfo", K_MSG, M_SHAPEINFO, false}, {4, "basisFuntions", K_MSG, M_BASIS, true}},
		/* M_SHAPEINFO proto:96-99  */ {{1, "functionShapeInfo", K_MSG, M_FTEST, false}, {2, "tetMesh", K_MSG, M_TETMESH, false}},
		/* M_FTEST     proto:19-21  */ {{1, "val", K_DOUBLE, M_COUNT, false}},
		/* M_TETMESH   proto:87-94  */ {{1, "vertices", K_MSG, M_VEC3, true}},
		/* M_VEC3      proto:30-34  */ {{1, "x", K_DOUBLE, M_COUNT, false}, {2, "y", K_DOUBLE, M_COUNT, false}, {3, "z", K_DOUBLE, M_COUNT, false}},
		/* M_BASIS     proto:101-104*/ {{1, "linearBspline", K_MSG, M_BSPLINE, false}, {2, "cubicBspline", K_MSG, M_BSPLINE, false}},
		/* M_BSPLINE   proto:106-116*/ {{1, "center", K_DOUBLE, M_COUNT, true}, {2, "support", K_DOUBLE, M_COUNT, true}, {3, "weight", K_DOUBLE, M_COUNT, false}},
		/* M_CELL      proto:118-122*/ {{1, "ranges", K_DOUBLE, M_COUNT, true}, {2, "leaf", K_MSG, M_LEAF, false}, {3, "interNode", K_MSG, M_INTER, false}},
		/* M_LEAF      proto:124-128*/ {{1, "centerShapeInfo", K_MSG, M_SHAPEINFO, false}, {2, "homeomorphicSamples", K_MSG, M_HSAMPLE, true}},
		/* M_HSAMPLE   proto:130-133*/ {{1, "precomputedSampleID", K_INT32, M_COUNT, false}, {2, "shapeInfo", K_MSG, M_SHAPEINFO, false}},
		/* M_INTER     proto:135-141*/ {{1, "child1", K_MSG, M_CELL, false}, {2, "child2", K_MSG, M_CELL, false}, {3, "splidDirection", K_INT32, M_COUNT, false}, {4, "splitVal", K_DOUBLE, M_COUNT, false}},
	};
	return S[t];
}
const FieldDef *find_field(MsgType t, int num)
{
	for (const FieldDef &f : schema(t)) if (f.num == num) return &f;
	return nullptr;
}
const FieldDef *find_field(MsgType t, const std::string &name)
{
	for (const FieldDef &f : schema(t)) if (name == f.name) return &f;
	return nullptr;
}

// ---- DOM ---------------------------------------------------------------------------------------------
struct Node;
struct Field {
	int num = 0;
	Kind kind = K_INT32;
	int64_t i = 0;
	double d = 0;
	std::unique_ptr<Node> m;
};
struct Node {
	MsgType type;
	std::vector<Field> f;
	explicit Node(MsgType t) : type(t) {}
	const Field *first(int num) const { for (const Field &x : f) if (x.num == num) return &x; return nullptr; }
	// proto2 "last one wins" for optional scalars / merge for messages is not needed for files the
	// reference writes (each optional field appears once); the first occurrence is used.
	void add_int(int num, int64_t v) { Field x; x.num = num; x.kind = K_INT32; x.i = v; f.push_back(std::move(x)); }
	void add_double(int num, double v) { Field x; x.num = num; x.kind = K_DOUBLE; x.d = v; f.push_back(std::move(x)); }
	Node *add_msg(int num, MsgType t) { Field x; x.num = num; x.kind = K_MSG; x.m.reset(new Node(t)); f.push_back(std::move(x)); return f.back().m.get(); }
};

// ---- binary decode -----------------------------------------------------------------------------------
struct Cursor { const uint8_t *p, *end; };

bool read_varint(Cursor &c, uint64_t *v)
{
	uint64_t r = 0;
	for (int shift = 0; shift < 64; shift += 7) {
		if (c.p >= c.end) return false;
		uint8_t b = *c.p++;
		r |= (uint64_t)(b & 0x7f) << shift;
		if (!(b & 0x80)) { *v = r; return true; }
	}
	return false;
}
bool read_fixed64(Cursor &c, double *d)
{
	if (c.end - c.p < 8) return false;
	memcpy(d, c.p, 8);  // little-endian host assumed (x86-64 / aarch64)
	c.p += 8;
	return true;
}

// protobuf's default recursion limit is 100 nested messages (SURVEY.md hazard H9); a tree level
// costs two (cell + interNode), so files the reference can load have depth <= ~48.  We allow more.
const int kMaxDepth = 4096;

bool decode_msg(Cursor c, Node *node, int depth)
{
	if (depth > kMaxDepth) return false;
	while (c.p < c.end) {
		uint64_t key;
		if (!read_varint(c, &key)) return false;
		int num = (int)(key >> 3), wt = (int)(key & 7);
		const FieldDef *fd = find_field(node->type, num);
		switch (wt) {
		case 0: {
			uint64_t v;
			if (!read_varint(c, &v)) return false;
			if (fd && fd->kind == K_INT32) node->add_int(num, (int64_t)(int32_t)(uint32_t)v);
			break;
		}
		case 1: {
			double d;
			if (!read_fixed64(c, &d)) return false;
			if (fd && fd->kind == K_DOUBLE) node->add_double(num, d);
			break;
		}
		case 2: {
			uint64_t len;
			if (!read_varint(c, &len)) return false;
			if ((uint64_t)(c.end - c.p) < len) return false;
			Cursor sub{c.p, c.p + len};
			c.p += len;
			if (!fd) break;
			if (fd->kind == K_MSG) {
				if (!decode_msg(sub, node->add_msg(num, fd->sub), depth + 1)) return false;
			} else if (fd->kind == K_DOUBLE) {  // packed repeated double (proto3-style writers)
				if (len % 8) return false;
				while (sub.p < sub.end) { double d; read_fixed64(sub, &d); node->add_double(num, d); }
			} else {                           // packed repeated int32
				while (sub.p < sub.end) { uint64_t v; if (!read_varint(sub, &v)) return false; node->add_int(num, (int64_t)(int32_t)(uint32_t)v); }
			}
			break;
		}
		case 5:
			if (c.end - c.p < 4) return false;
			c.p += 4;
			break;
		default:
			return false;  // groups (3,4) are not used by this schema
		}
	}
	return true;
}

// ---- binary encode ----------------------------------------------------------------------------------------
void put_varint(std::string &o, uint64_t v)
{
	while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
	o.push_back((char)v);
}
void encode_msg(const Node &n, std::string &o)
{
	// protobuf's generated serializers emit fields in field-number order; the DOM is built that way.
	for (const Field &x : n.f) {
		switch (x.kind) {
		case K_INT32:
			put_varint(o, (uint64_t)x.num << 3 | 0);
			put_varint(o, (uint64_t)(int64_t)(int32_t)x.i);  // negative int32 -> 10-byte varint, as protobuf does
			break;
		case K_DOUBLE: {
			put_varint(o, (uint64_t)x.num << 3 | 1);  // proto2 repeated doubles are NOT packed
			char b[8];
			memcpy(b, &x.d, 8);
			o.append(b, 8);
			break;
		}
		case K_MSG: {
			std::string sub;
			encode_msg(*x.m, sub);
			put_varint(o, (uint64_t)x.num << 3 | 2);
			put_varint(o, sub.size());
			o += sub;
			break;
		}
		}
	}
}

// ---- text format ----------------------------------------------------------------------------------------------
std::string dtoa(double v)
{
	if (std::isnan(v)) return "nan";
	if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
	char buf[40];
	snprintf(buf, sizeof buf, "%.15g", v);  // protobuf SimpleDtoa: 15 digits if they round-trip, else 17
	if (strtod(buf, nullptr) != v) snprintf(buf, sizeof buf, "%.17g", v);
	return buf;
}
void text_msg(const Node &n, std::string &o, int indent)
{
	std::string pad((size_t)indent * 2, ' ');
	for (const Field &x : n.f) {
		const FieldDef *fd = find_field(n.type, x.num);
		switch (x.kind) {
		case K_INT32: o += pad + fd->name + ": " + std::to_string(x.i) + "\n"; break;
		case K_DOUBLE: o += pad + fd->name + ": " + dtoa(x.d) + "\n"; break;
		case K_MSG:
			o += pad + fd->name + " {\n";
			text_msg(*x.m, o, indent + 1);
			o += pad + "}\n";
			break;
		}
	}
}

struct Lexer {
	const char *p, *end;
	void skip()
	{
		for (;;) {
			while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == ',' || *p == ';')) p++;
			if (p < end && *p == '#') { while (p < end && *p != '\n') p++; continue; }
			break;
		}
	}
	bool eof() { skip(); return p >= end; }
	bool accept(char c) { skip(); if (p < end && *p == c) { p++; return true; } return false; }
	std::string word()
	{
		skip();
		const char *s = p;
		while (p < end && (isalnum((unsigned char)*p) || *p == '_' || *p == '.' || *p == '-' || *p == '+')) p++;
		return std::string(s, p);
	}
};

bool parse_text_value_skip(Lexer &lx);
bool parse_text_msg(Lexer &lx, Node *node, char closer, int depth)
{
	if (depth > kMaxDepth) return false;
	for (;;) {
		if (closer ? lx.accept(closer) : lx.eof()) return true;
		if (lx.eof()) return false;
		std::string name = lx.word();
		if (name.empty()) return false;
		const FieldDef *fd = find_field(node->type, name);
		bool colon = lx.accept(':');
		char close = 0;
		if (lx.accept('{')) close = '}';
		else if (lx.accept('<')) close = '>';
		if (close) {
			if (fd && fd->kind == K_MSG) {
				if (!parse_text_msg(lx, node->add_msg(fd->num, fd->sub), close, depth + 1)) return false;
			} else {
				Node scratch(M_VEC3);  // unknown sub-message: parse and drop
				if (!parse_text_msg(lx, &scratch, close, depth + 1)) return false;
			}
			continue;
		}
		if (!colon) return false;
		lx.skip();
		if (lx.p < lx.end && (*lx.p == '"' || *lx.p == '\'')) {  // string value of a field we do not use
			char q = *lx.p++;
			while (lx.p < lx.end && *lx.p != q) { if (*lx.p == '\\') lx.p++; lx.p++; }
			if (lx.p >= lx.end) return false;
			lx.p++;
			continue;
		}
		std::string tok = lx.word();
		if (tok.empty()) return false;
		if (!fd) continue;
		if (fd->kind == K_INT32) node->add_int(fd->num, strtoll(tok.c_str(), nullptr, 0));
		else if (fd->kind == K_DOUBLE) {
			std::string t = tok;
			if (!t.empty() && (t.back() == 'f' || t.back() == 'F') && t != "inf" && t != "-inf") t.pop_back();
			node->add_double(fd->num, strtod(t.c_str(), nullptr));
		} else return false;
	}
}

// ---- DOM -> Graph  (convertToCplusplus, LCProtoConverter.cpp:356-403) -------------------------------------
struct Loader {
	Graph *g;
	std::unordered_map<int64_t, int32_t> id_to_sample;
	bool have_type = false;

	// convertShapeInfoToCpp (LCProtoConverter.cpp:265-277); tetMesh.vertices is the mesh-valued slot
	// the public reference removed (:211-263) -- D = 3 * #vertices.
	Status value_of(const Node &si, std::vector<double> *v)
	{
		v->clear();
		if (const Field *f = si.first(1)) {
			const Field *val = f->m->first(1);
			if (!val) return Status("value for functio shape info not found");
			v->push_back(val->d);
			return Status();
		}
		if (const Field *t = si.first(2)) {
			for (const Field &vx : t->m->f) {
				if (vx.num != 1) continue;
				const Field *x = vx.m->first(1), *y = vx.m->first(2), *z = vx.m->first(3);
				v->push_back(x ? x->d : 0.0);
				v->push_back(y ? y->d : 0.0);
				v->push_back(z ? z->d : 0.0);
			}
			if (!v->empty()) return Status();
		}
		return Status("type fo shape info not recognized");
	}
	Status add_row(const std::vector<double> &v, int32_t *row)
	{
		if (g->values.empty() && g->sample_row.empty() && g->tree.size() == 0) g->D = (int32_t)v.size();
		if ((int32_t)v.size() != g->D) return Status("shape info values of different sizes in one file");
		*row = (int32_t)g->n_rows();
		g->values.insert(g->values.end(), v.begin(), v.end());
		return Status();
	}
	// per-cell copy: reuse the sample's own row when the copy is bit-identical (always true for
	// LCRealFunctionValue::getHomeophicShapeInfo, LCRealFunctionValue.cpp:63-73)
	Status copy_row(const std::vector<double> &v, int32_t sample, int32_t *row)
	{
		int32_t sr = g->sample_row[sample];
		if ((int32_t)v.size() == g->D && memcmp(v.data(), &g->values[(size_t)sr * g->D], sizeof(double) * v.size()) == 0) { *row = sr; return Status(); }
		return add_row(v, row);
	}

	// convertBasisFuntionToCpp (LCProtoConverter.cpp:172-209)
	Status basis(const Node &bf)
	{
		const Field *lin = bf.first(1), *cub = bf.first(2);
		const Field *src = lin ? lin : cub;
		if (!src) return Status("type of basis function not recognized");
		int32_t type = lin ? LINEAR_BSPLINE : CUBIC_BSPLINE;
		if (!have_type) { g->basis = type; have_type = true; }  // LCSampledFunction.cpp:15
		else if (type != g->basis) return Status("mixed basis function types in one function are not supported");
		int nc = 0, ns = 0;
		double w = 0;
		for (const Field &x : src->m->f) {
			if (x.num == 1) { g->bf_center.push_back(x.d); nc++; }
			else if (x.num == 2) { g->bf_support.push_back(x.d); ns++; }
			else if (x.num == 3) w = x.d;
		}
		if (nc != g->N || ns != g->N) return Status("basis function center/support of wrong size");
		g->bf_weight.push_back(w);
		return Status();
	}

	// convertPrecomputedSampleToCpp (LCProtoConverter.cpp:279-306)
	Status sample(const Node &s)
	{
		const Field *id = s.first(1);
		if (!id) return Status("sample does not have id");
		int nc = 0;
		for (const Field &x : s.f) if (x.num == 2) { g->sample_center.push_back(x.d); nc++; }
		if (nc != g->N) return Status("sample center of wrong size");
		const Field *si = s.first(3);
		if (!si) return Status("sample does not have shape info");
		std::vector<double> v;
		LOCO_RETURN_IF_ERROR(value_of(*si->m, &v));
		int32_t row;
		LOCO_RETURN_IF_ERROR(add_row(v, &row));
		int32_t index = (int32_t)g->sample_row.size();
		g->sample_row.push_back(row);
		for (const Field &x : s.f) if (x.num == 4) LOCO_RETURN_IF_ERROR(basis(*x.m));
		g->bf_off.push_back((int64_t)g->bf_weight.size());
		id_to_sample[id->i] = index;  // idsToSamples[id] = sample (later ids overwrite earlier, as in the map)
		return Status();
	}

	// convertetAdaptiveGridCellToCpp (LCProtoConverter.cpp:308-354); returns the cell's index
	Status cell(const Node &c, Graph::CellSet *cs, int depth)
	{
		if (depth > kMaxDepth) return Status("tree too deep");
		int nr = 0;
		for (const Field &x : c.f) if (x.num == 1) { cs->ranges.push_back(x.d); nr++; }
		if (nr != 2 * g->N) return Status("cell ranges of wrong size");
		size_t me = cs->split_dir.size();
		cs->split_dir.push_back(-1);
		cs->split_val.push_back(0.0);
		cs->child2.push_back(-1);
		cs->center_row.push_back(-1);
		if (const Field *leaf = c.first(2)) {
			std::vector<double> v;
			const Field *csi = leaf->m->first(1);
			if (!csi) return Status("type fo shape info not recognized");
			LOCO_RETURN_IF_ERROR(value_of(*csi->m, &v));
			int32_t row;
			LOCO_RETURN_IF_ERROR(add_row(v, &row));
			cs->center_row[me] = row;
			// samples_ is a map keyed by sample: a repeated id keeps one entry, the last value wins
			std::unordered_map<int32_t, size_t> seen;
			for (const Field &h : leaf->m->f) {
				if (h.num != 2) continue;
				const Field *sid = h.m->first(1);
				auto it = id_to_sample.find(sid ? sid->i : 0);
				if (it == id_to_sample.end()) return Status("could not map the sample from id");
				const Field *hsi = h.m->first(2);
				if (!hsi) return Status("type fo shape info not recognized");
				LOCO_RETURN_IF_ERROR(value_of(*hsi->m, &v));
				int32_t r;
				LOCO_RETURN_IF_ERROR(copy_row(v, it->second, &r));
				auto dup = seen.find(it->second);
				if (dup != seen.end()) { cs->ls_row[dup->second] = r; continue; }
				seen[it->second] = cs->ls_sample.size();
				cs->ls_sample.push_back(it->second);
				cs->ls_row.push_back(r);
			}
			cs->ls_off.push_back((int64_t)cs->ls_sample.size());
			return Status();
		}
		if (const Field *in = c.first(3)) {
			const Field *c1 = in->m->first(1), *c2 = in->m->first(2), *sd = in->m->first(3), *sv = in->m->first(4);
			if (!c1 || !c2) return Status("not an interal node or ");
			cs->split_dir[me] = sd ? (int32_t)sd->i : 0;
			cs->split_val[me] = sv ? sv->d : 0.0;
			if (cs->split_dir[me] < 0 || cs->split_dir[me] >= g->N) return Status("split direction out of range");
			cs->ls_off.push_back((int64_t)cs->ls_sample.size());
			LOCO_RETURN_IF_ERROR(cell(*c1->m, cs, depth + 1));
			cs->child2[me] = (int32_t)cs->split_dir.size();
			LOCO_RETURN_IF_ERROR(cell(*c2->m, cs, depth + 1));
			return Status();
		}
		return Status("not an interal node or ");
	}

	Status run(const Node &root)
	{
		*g = Graph();
		const Field *rc = root.first(1);
		if (!rc) return Status("not an interal node or ");
		int nr = 0;
		for (const Field &x : rc->m->f) if (x.num == 1) nr++;
		if (nr < 2 || nr % 2) return Status("cell ranges of wrong size");
		g->N = nr / 2;
		g->bf_off.push_back(0);
		g->tree.ls_off.push_back(0);
		g->border.ls_off.push_back(0);
		// step 1: all precomputed samples (:361-370)
		for (const Field &x : root.f) if (x.num == 2) LOCO_RETURN_IF_ERROR(sample(*x.m));
		if (g->sample_row.empty()) return Status("function has no samples");
		if (g->bf_off[1] == 0) return Status("sample 0 has no basis function (basis type undefined)");
		// root, then border cells loaded as parentless level-0 leaves (:371-386)
		LOCO_RETURN_IF_ERROR(cell(*rc->m, &g->tree, 0));
		for (const Field &x : root.f) {
			if (x.num != 4) continue;
			size_t before = g->border.split_dir.size();
			LOCO_RETURN_IF_ERROR(cell(*x.m, &g->border, 0));
			if (g->border.split_dir.size() != before + 1 || g->border.split_dir[before] != -1) return Status("border cells must be leaves");
		}
		// ranges default to the root's (:388-396)
		for (const Field &x : root.f) if (x.num == 5) g->domain.push_back(x.d);
		if (g->domain.empty()) g->domain.assign(g->tree.ranges.begin(), g->tree.ranges.begin() + 2 * g->N);
		if ((int)g->domain.size() != 2 * g->N) return Status("range of wrong size");
		return validate_graph(*g);
	}
};

// ---- Graph -> DOM  (convertToProto, LCProtoConverter.cpp:141-168) ------------------------------------------------
Status put_value(const Graph &g, int32_t row, Node *si)
{
	if (row < 0 || row >= g.n_rows()) return Status("undefined type of shape info");
	const double *v = &g.values[(size_t)row * g.D];
	if (g.D == 1) { si->add_msg(1, M_FTEST)->add_double(1, v[0]); return Status(); }
	if (g.D % 3) return Status("value dimension must be 1 or a multiple of 3 (mesh vertices) to be stored");
	Node *tm = si->add_msg(2, M_TETMESH);
	for (int j = 0; j < g.D; j += 3) {
		Node *vx = tm->add_msg(1, M_VEC3);
		vx->add_double(1, v[j]); vx->add_double(2, v[j + 1]); vx->add_double(3, v[j + 2]);
	}
	return Status();
}
Status put_cell(const Graph &g, const Graph::CellSet &cs, int32_t i, Node *c, int32_t *next)
{
	for (int d = 0; d < 2 * g.N; d++) c->add_double(1, cs.ranges[(size_t)i * 2 * g.N + d]);
	if (cs.split_dir[i] < 0) {
		Node *leaf = c->add_msg(2, M_LEAF);
		LOCO_RETURN_IF_ERROR(put_value(g, cs.center_row[i], leaf->add_msg(1, M_SHAPEINFO)));
		for (int64_t e = cs.ls_off[i]; e < cs.ls_off[i + 1]; e++) {
			Node *h = leaf->add_msg(2, M_HSAMPLE);
			h->add_int(1, cs.ls_sample[e]);
			LOCO_RETURN_IF_ERROR(put_value(g, cs.ls_row[e], h->add_msg(2, M_SHAPEINFO)));
		}
		*next = i + 1;
		return Status();
	}
	Node *in = c->add_msg(3, M_INTER);
	int32_t after1, after2;
	LOCO_RETURN_IF_ERROR(put_cell(g, cs, i + 1, in->add_msg(1, M_CELL), &after1));
	LOCO_RETURN_IF_ERROR(put_cell(g, cs, cs.child2[i], in->add_msg(2, M_CELL), &after2));
	in->add_int(3, cs.split_dir[i]);
	in->add_double(4, cs.split_val[i]);
	*next = after2;
	return Status();
}
Status graph_to_dom(const Graph &g, Node *root)
{
	LOCO_RETURN_IF_ERROR(validate_graph(g));
	int32_t next;
	LOCO_RETURN_IF_ERROR(put_cell(g, g.tree, 0, root->add_msg(1, M_CELL), &next));
	for (int64_t s = 0; s < g.n_samples(); s++) {
		Node *ps = root->add_msg(2, M_SAMPLE);
		ps->add_int(1, s);  // id = index in samples_ (:147-152)
		for (int d = 0; d < g.N; d++) ps->add_double(2, g.sample_center[(size_t)s * g.N + d]);
		LOCO_RETURN_IF_ERROR(put_value(g, g.sample_row[s], ps->add_msg(3, M_SHAPEINFO)));
		for (int64_t b = g.bf_off[s]; b < g.bf_off[s + 1]; b++) {
			Node *bs = ps->add_msg(4, M_BASIS)->add_msg(g.basis == LINEAR_BSPLINE ? 1 : 2, M_BSPLINE);
			for (int d = 0; d < g.N; d++) bs->add_double(1, g.bf_center[(size_t)b * g.N + d]);
			for (int d = 0; d < g.N; d++) bs->add_double(2, g.bf_support[(size_t)b * g.N + d]);
			bs->add_double(3, g.bf_weight[b]);
		}
	}
	for (int32_t b = 0; b < g.border.size(); b++)
		LOCO_RETURN_IF_ERROR(put_cell(g, g.border, b, root->add_msg(4, M_CELL), &next));
	for (double r : g.domain) root->add_double(5, r);
	return Status();
}

} // namespace

// ---- public ----------------------------------------------------------------------------------------------------------
Status validate_graph(const Graph &g)
{
	if (g.N < 1) return Status("invalid number of parameters");
	if (g.D < 1) return Status("invalid value dimension");
	if (g.basis != LINEAR_BSPLINE && g.basis != CUBIC_BSPLINE) return Status("basis type unknown");
	int64_t S = g.n_samples(), R = g.n_rows();
	if ((int64_t)g.domain.size() != 2 * g.N) return Status("range of wrong size");
	if ((int64_t)g.sample_center.size() != S * g.N || (int64_t)g.bf_off.size() != S + 1) return Status("sample arrays of wrong size");
	int64_t B = g.bf_off.empty() ? 0 : g.bf_off.back();
	if ((int64_t)g.bf_weight.size() != B || (int64_t)g.bf_center.size() != B * g.N || (int64_t)g.bf_support.size() != B * g.N) return Status("basis function arrays of wrong size");
	for (int64_t s = 0; s < S; s++) if (g.sample_row[s] < 0 || g.sample_row[s] >= R) return Status("sample value row out of range");
	for (const Graph::CellSet *cs : {&g.tree, &g.border}) {
		int64_t C = cs->size();
		if ((int64_t)cs->ranges.size() != C * 2 * g.N || (int64_t)cs->ls_off.size() != C + 1 || (int64_t)cs->child2.size() != C ||
		    (int64_t)cs->split_val.size() != C || (int64_t)cs->center_row.size() != C) return Status("cell arrays of wrong size");
		for (int64_t i = 0; i < C; i++) {
			if (cs->split_dir[i] >= g.N) return Status("split direction out of range");
			if (cs->split_dir[i] >= 0 && (cs->child2[i] <= i + 1 || cs->child2[i] >= C)) return Status("child index out of range");
		}
		for (size_t e = 0; e < cs->ls_sample.size(); e++) {
			if (cs->ls_sample[e] < 0 || cs->ls_sample[e] >= S) return Status("could not map the sample from id");
			if (cs->ls_row[e] < 0 || cs->ls_row[e] >= R) return Status("cell value row out of range");
		}
	}
	if (g.tree.size() < 1) return Status("function has no root cell");
	return Status();
}

Status decode_proto(const uint8_t *data, size_t size, Graph *out)
{
	Node root(M_SHAPE);
	if (!decode_msg(Cursor{data, data + size}, &root, 0)) return Status("could not parse the file");
	Loader ld;
	ld.g = out;
	return ld.run(root);
}

Status decode_proto_text(const std::string &text, Graph *out)
{
	Node root(M_SHAPE);
	Lexer lx{text.data(), text.data() + text.size()};
	if (!parse_text_msg(lx, &root, 0, 0)) return Status("could not parse the file");
	Loader ld;
	ld.g = out;
	return ld.run(root);
}

Status encode_proto(const Graph &g, std::string *out)
{
	Node root(M_SHAPE);
	LOCO_RETURN_IF_ERROR(graph_to_dom(g, &root));
	out->clear();
	encode_msg(root, *out);
	return Status();
}

Status encode_proto_text(const Graph &g, std::string *out)
{
	Node root(M_SHAPE);
	LOCO_RETURN_IF_ERROR(graph_to_dom(g, &root));
	out->clear();
	text_msg(root, *out, 0);
	return Status();
}

// ProtoConverter<T>::loadFromFile (LCShapeInfoProtoConverter.h:46-77)
Status read_proto_file(const std::string &path, bool ascii, Graph *out)
{
	std::ifstream in(path, ascii ? std::ios::in : (std::ios::in | std::ios::binary));
	if (!in) return Status("could not open file " + path);
	std::ostringstream ss;
	ss << in.rdbuf();
	std::string buf = ss.str();
	return ascii ? decode_proto_text(buf, out) : decode_proto((const uint8_t *)buf.data(), buf.size(), out);
}

// ProtoConverter<T>::saveToFile (LCShapeInfoProtoConverter.h:15-43)
Status write_proto_file(const Graph &g, const std::string &path, bool ascii)
{
	std::string buf;
	LOCO_RETURN_IF_ERROR(ascii ? encode_proto_text(g, &buf) : encode_proto(g, &buf));
	std::ofstream outf(path, ascii ? std::ios::out : (std::ios::out | std::ios::binary));
	if (!outf) return Status("could not write to file " + path);
	outf.write(buf.data(), (std::streamsize)buf.size());
	return Status();
}

} // namespace loco
