// shim/gflags/gflags.h -- the subset of gflags the reference's examples use (DEFINE_string / int32 / bool / double,
// gflags::ParseCommandLineFlags), so that examples/*.cpp compile unchanged where gflags is not installed.
// Accepted syntax: --name=value, --name value, --flag / --noflag / --flag=true|false for booleans, "--" ends the flags;
// an unknown flag or a malformed value is an error (exit 1), like gflags.  With remove_flags the parsed arguments are
// removed from argv.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

namespace gflags {
namespace shim {
    struct Flag {
        enum Kind { STRING, INT32, BOOL, DOUBLE } kind;
        void* ptr;
        const char* help;
    };
    inline std::map<std::string, Flag>& registry()
    {
        static std::map<std::string, Flag> r;
        return r;
    }
    struct Registrar {
        Registrar(const char* name, Flag::Kind k, void* p, const char* help) { registry()[name] = Flag{ k, p, help }; }
    };
    inline bool set(const Flag& f, const std::string& v)
    {
        char* end = nullptr;
        switch (f.kind) {
        case Flag::STRING: *static_cast<std::string*>(f.ptr) = v; return true;
        case Flag::INT32: { const long x = std::strtol(v.c_str(), &end, 0); if (v.empty() || *end) return false; *static_cast<int32_t*>(f.ptr) = (int32_t)x; return true; }
        case Flag::DOUBLE: { const double x = std::strtod(v.c_str(), &end); if (v.empty() || *end) return false; *static_cast<double*>(f.ptr) = x; return true; }
        case Flag::BOOL:
            if (v == "true" || v == "1" || v == "t" || v == "yes" || v == "y") { *static_cast<bool*>(f.ptr) = true; return true; }
            if (v == "false" || v == "0" || v == "f" || v == "no" || v == "n") { *static_cast<bool*>(f.ptr) = false; return true; }
            return false;
        }
        return false;
    }
}

inline uint32_t ParseCommandLineFlags(int* argc, char*** argv, bool remove_flags)
{
    std::vector<char*> keep{ (*argv)[0] };
    int i = 1;
    for (; i < *argc; ++i) {
        const char* a = (*argv)[i];
        if (std::strcmp(a, "--") == 0) { ++i; break; }
        if (a[0] != '-') { keep.push_back((*argv)[i]); continue; }
        std::string body(a + ((a[1] == '-') ? 2 : 1));
        std::string name = body, value;
        bool has_value = false;
        const size_t eq = body.find('=');
        if (eq != std::string::npos) { name = body.substr(0, eq); value = body.substr(eq + 1); has_value = true; }
        auto& reg = shim::registry();
        auto it = reg.find(name);
        if (it == reg.end() && !has_value && name.rfind("no", 0) == 0) {   // --noflag
            auto it2 = reg.find(name.substr(2));
            if (it2 != reg.end() && it2->second.kind == shim::Flag::BOOL) { *static_cast<bool*>(it2->second.ptr) = false; continue; }
        }
        if (it == reg.end()) { std::cerr << "ERROR: unknown command line flag '" << name << "'\n"; std::exit(1); }
        if (!has_value) {
            if (it->second.kind == shim::Flag::BOOL) { *static_cast<bool*>(it->second.ptr) = true; continue; }
            if (i + 1 >= *argc) { std::cerr << "ERROR: flag '" << name << "' is missing its argument\n"; std::exit(1); }
            value = (*argv)[++i];
        }
        if (!shim::set(it->second, value)) { std::cerr << "ERROR: illegal value '" << value << "' specified for flag '" << name << "'\n"; std::exit(1); }
    }
    for (; i < *argc; ++i) keep.push_back((*argv)[i]);
    if (remove_flags) {
        for (size_t k = 0; k < keep.size(); ++k) (*argv)[k] = keep[k];
        *argc = (int)keep.size();
    }
    return remove_flags ? 1u : (uint32_t)i;
}
} // namespace gflags
namespace google = gflags;

#define HPB_GFLAGS_DEFINE(type, kind, name, val, txt)                                            \
    type FLAGS_##name = val;                                                                     \
    static ::gflags::shim::Registrar hpb_flag_registrar_##name(#name, ::gflags::shim::Flag::kind, &FLAGS_##name, txt)
#define DEFINE_string(name, val, txt) HPB_GFLAGS_DEFINE(std::string, STRING, name, val, txt)
#define DEFINE_int32(name, val, txt) HPB_GFLAGS_DEFINE(int32_t, INT32, name, val, txt)
#define DEFINE_bool(name, val, txt) HPB_GFLAGS_DEFINE(bool, BOOL, name, val, txt)
#define DEFINE_double(name, val, txt) HPB_GFLAGS_DEFINE(double, DOUBLE, name, val, txt)
