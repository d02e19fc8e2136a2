// Minimal host for Qt's ECMAScript engine (QJSEngine, libQt6Qml 6.6.3 as shipped inside
// Nsight Compute's host directory) -- the only JavaScript engine in this image.
// No Qt headers are installed, so the handful of classes used are declared here with the
// layout/ABI of Qt 6 (QString = QArrayDataPointer{d,ptr,size}; QJSValue = one word;
// QObject = vptr + d_ptr).  Only out-of-line exported members are called.
//
// usage: qjs_run file1.js [file2.js ...]   -> evaluates the files in order in one engine,
//        prints the string value of the LAST file's completion value on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct QByteArrayView { long long size; const char* data; };
class QByteArray {
public:
    void* d; char* ptr; long long size;
    ~QByteArray() {}   // inline in Qt (ref-count drop); leaking is fine for a one-shot tool
};
class QString {
public:
    void* d; char16_t* ptr; long long size;
    QString() : d(nullptr), ptr(nullptr), size(0) {}
    ~QString() {}
    static QString fromUtf8(QByteArrayView);
    QByteArray toUtf8() const { return toUtf8_helper(*this); }
    static QByteArray toUtf8_helper(const QString&);
};
template <class T> class QList;
class QObject { public: virtual ~QObject(); void* d_ptr; };
class QCoreApplication : public QObject {
public:
    QCoreApplication(int& argc, char** argv, int flags = 0x060603 /* QT_VERSION 6.6.3 */);
    ~QCoreApplication() override;
};
class QJSValue {
public:
    unsigned long long d;
    ~QJSValue();
    bool isError() const;
    QString toString() const;
    QJSValue property(const QString&) const;
};
class QJSEngine : public QObject {
public:
    QJSEngine();
    ~QJSEngine() override;
    QJSValue evaluate(const QString& program, const QString& fileName, int lineNumber, QList<QString>* stack);
};

static std::string slurp(const char* p) {
    FILE* f = fopen(p, "rb"); if (!f) { fprintf(stderr, "cannot open %s\n", p); exit(2); }
    std::string s; char buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f); return s;
}
static QString qs(const std::string& s) { return QString::fromUtf8(QByteArrayView{(long long)s.size(), s.data()}); }
static std::string str(const QString& q) { QByteArray b = q.toUtf8(); return std::string(b.ptr ? b.ptr : "", (size_t)b.size); }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: qjs_run a.js [b.js ...]\n"); return 2; }
    QCoreApplication app(argc, argv);
    QJSEngine eng;
    std::string last;
    for (int i = 1; i < argc; i++) {
        QJSValue v = eng.evaluate(qs(slurp(argv[i])), qs(argv[i]), 1, nullptr);
        if (v.isError()) {
            fprintf(stderr, "%s: JS error: %s (line %s)\n", argv[i], str(v.toString()).c_str(),
                    str(v.property(qs("lineNumber")).toString()).c_str());
            return 1;
        }
        last = str(v.toString());
    }
    fwrite(last.data(), 1, last.size(), stdout);
    fputc('\n', stdout);
    return 0;
}
