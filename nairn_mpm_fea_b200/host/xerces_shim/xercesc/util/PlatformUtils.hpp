// TEST INFRASTRUCTURE (oracle build only) -- not product code.
//
// Minimal stand-in for the Xerces-C SAX2 surface that nairn-mpm-fea's XML
// reader uses, implemented over expat (which this image has; Xerces-C is not
// installed).  It exists only so that oracle/build_ref.sh can compile the
// reference's own sources, unmodified and where they lie under
// /root/reference, into oracle/_ref/.  Surface taken from
// Common/Read_XML/CommonReadHandler.hpp:25-33 and
// Common/System/CommonAnalysis.cpp:184-273.
//
// Behavioural requirement honoured here: character data is coalesced and
// delivered once before the next start/end tag (the reference's
// CommonReadHandler::characters applies and resets its scaling on first call).
#ifndef MPMGPU_XERCES_SHIM_H
#define MPMGPU_XERCES_SHIM_H

#include <expat.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define XERCES_CPP_NAMESPACE_USE
typedef char16_t XMLCh;          // must differ from char so transcode() overloads resolve
typedef size_t XMLSize_t;
typedef unsigned long XMLFileLoc;

typedef std::basic_string<XMLCh> XShimString;

inline XShimString XShimWiden(const char *s)
{
    XShimString r;
    for (; s && *s; ++s) r.push_back((XMLCh)(unsigned char)*s);
    return r;
}

class XMLString {
public:
    static char *transcode(const XMLCh *s)
    {
        size_t n = 0;
        while (s && s[n]) n++;
        char *r = new char[n + 1];
        for (size_t i = 0; i < n; i++) r[i] = (char)s[i];
        r[n] = 0;
        return r;
    }
    static XMLCh *transcode(const char *s)
    {
        size_t n = strlen(s);
        XMLCh *r = new XMLCh[n + 1];
        for (size_t i = 0; i < n; i++) r[i] = (XMLCh)(unsigned char)s[i];
        r[n] = 0;
        return r;
    }
    static void release(char **p) { delete[] *p; *p = NULL; }
    static void release(XMLCh **p) { delete[] *p; *p = NULL; }
};

class XMLException {
    XShimString msg;
public:
    XMLException(const char *m) : msg(XShimWiden(m)) {}
    const XMLCh *getMessage() const { return msg.c_str(); }
};

class SAXException {
    XShimString msg;
public:
    SAXException(const char *m) : msg(XShimWiden(m)) {}
    SAXException(const XMLCh *m) : msg(m) {}
    virtual ~SAXException() {}
    const XMLCh *getMessage() const { return msg.c_str(); }
};

class SAXParseException : public SAXException {
    XShimString sys;
    XMLFileLoc line, col;
public:
    SAXParseException(const char *m, const char *s, XMLFileLoc l, XMLFileLoc c)
        : SAXException(m), sys(XShimWiden(s)), line(l), col(c) {}
    const XMLCh *getSystemId() const { return sys.c_str(); }
    XMLFileLoc getLineNumber() const { return line; }
    XMLFileLoc getColumnNumber() const { return col; }
};

class Attributes {
public:
    std::vector<XShimString> names, values;
    XMLSize_t getLength() const { return names.size(); }
    const XMLCh *getLocalName(XMLSize_t i) const { return names[i].c_str(); }
    const XMLCh *getValue(XMLSize_t i) const { return values[i].c_str(); }
};

class DefaultHandler {
public:
    virtual ~DefaultHandler() {}
    virtual void startElement(const XMLCh *const, const XMLCh *const, const XMLCh *const, const Attributes &) {}
    virtual void endElement(const XMLCh *const, const XMLCh *const, const XMLCh *const) {}
    virtual void characters(const XMLCh *const, const XMLSize_t) {}
    virtual void warning(const SAXParseException &) {}
    virtual void error(const SAXParseException &) {}
    virtual void fatalError(const SAXParseException &) {}
};

class SAX2XMLReader {
    DefaultHandler *handler;
    std::string pending;        // coalesced character data

    void flushText()
    {
        if (pending.empty()) return;
        XShimString x = XShimWiden(pending.c_str());
        pending.clear();
        handler->characters(x.c_str(), x.size());
    }
    static void XMLCALL onStart(void *u, const char *name, const char **atts)
    {
        SAX2XMLReader *r = (SAX2XMLReader *)u;
        r->flushText();
        Attributes a;
        for (int i = 0; atts[i]; i += 2) {
            a.names.push_back(XShimWiden(atts[i]));
            a.values.push_back(XShimWiden(atts[i + 1]));
        }
        XShimString nm = XShimWiden(name);
        r->handler->startElement(NULL, nm.c_str(), nm.c_str(), a);
    }
    static void XMLCALL onEnd(void *u, const char *name)
    {
        SAX2XMLReader *r = (SAX2XMLReader *)u;
        r->flushText();
        XShimString nm = XShimWiden(name);
        r->handler->endElement(NULL, nm.c_str(), nm.c_str());
    }
    static void XMLCALL onText(void *u, const char *s, int len)
    {
        ((SAX2XMLReader *)u)->pending.append(s, len);
    }

public:
    SAX2XMLReader() : handler(NULL) {}
    void setFeature(const XMLCh *, bool) {}
    void setContentHandler(DefaultHandler *h) { handler = h; }
    void setErrorHandler(DefaultHandler *) {}
    void parse(const char *file)
    {
        FILE *f = fopen(file, "rb");
        if (!f) throw SAXException("cannot open input file");
        XML_Parser p = XML_ParserCreate(NULL);
        XML_SetUserData(p, this);
        XML_SetElementHandler(p, onStart, onEnd);
        XML_SetCharacterDataHandler(p, onText);
        char buf[65536];
        size_t n;
        do {
            n = fread(buf, 1, sizeof buf, f);
            if (XML_Parse(p, buf, (int)n, n == 0) == XML_STATUS_ERROR) {
                SAXParseException e(XML_ErrorString(XML_GetErrorCode(p)), file,
                                    XML_GetCurrentLineNumber(p), XML_GetCurrentColumnNumber(p));
                fclose(f);
                XML_ParserFree(p);
                handler->fatalError(e);
                throw e;
            }
        } while (n > 0);
        fclose(f);
        XML_ParserFree(p);
    }
};

class XMLReaderFactory {
public:
    static SAX2XMLReader *createXMLReader() { return new SAX2XMLReader(); }
};

class XMLPlatformUtils {
public:
    static void Initialize() {}
    static void Terminate() {}
};

#endif
