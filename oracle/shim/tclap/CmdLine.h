// TEST INFRASTRUCTURE ONLY (oracle/): the TCLAP subset dftcxx.cpp uses.
#ifndef DFTB200_ORACLE_SHIM_TCLAP
#define DFTB200_ORACLE_SHIM_TCLAP
#include <string>
#include <vector>
namespace TCLAP {
class ArgException : public std::exception {
    std::string e_, id_;

public:
    ArgException(const std::string& e, const std::string& id) : e_(e), id_(id) {}
    std::string error() const { return e_; }
    std::string argId() const { return id_; }
    const char* what() const noexcept override { return e_.c_str(); }
};
template <typename T>
class ValueArg {
public:
    std::string flag, name;
    bool req, set;
    T value;
    ValueArg(const std::string& f, const std::string& n, const std::string&, bool r, const T& def, const std::string&)
        : flag(f), name(n), req(r), set(false), value(def) {}
    const T& getValue() const { return value; }
};
class CmdLine {
    std::vector<ValueArg<std::string>*> args;

public:
    CmdLine(const std::string&, char, const std::string&) {}
    void add(ValueArg<std::string>& a) { args.push_back(&a); }
    void parse(int argc, char** argv) {
        for (int i = 1; i < argc; i++) {
            std::string s(argv[i]);
            bool matched = false;
            for (auto* a : args)
                if (s == "-" + a->flag || s == "--" + a->name) {
                    if (i + 1 >= argc) throw ArgException("Missing a value for this argument!", a->name);
                    a->value = argv[++i];
                    a->set = true;
                    matched = true;
                }
            if (!matched) throw ArgException("Couldn't find match for argument", s);
        }
        for (auto* a : args)
            if (a->req && !a->set) throw ArgException("Required argument missing", a->name);
    }
};
}  // namespace TCLAP
#endif
