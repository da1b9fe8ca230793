/* libxml2 STUB -- test infrastructure only (oracle/).  libxml2 is absent; the
 * reference's src/xmlparser.h only needs these pointer typedefs to be declared
 * for src/loudspeakerrenderer.h to parse.  The harness adds outputs with
 * renderer.add(Output::Params()) and never calls load_reproduction_setup(),
 * so no XML function is ever instantiated or linked. */
#ifndef SSR_B200_ORACLE_LIBXML_STUB_H
#define SSR_B200_ORACLE_LIBXML_STUB_H
typedef struct ssr_b200_xml_doc_stub* xmlDocPtr;
typedef struct ssr_b200_xml_node_stub* xmlNodePtr;
typedef struct ssr_b200_xml_xpathctx_stub* xmlXPathContextPtr;
typedef struct ssr_b200_xml_xpathobj_stub* xmlXPathObjectPtr;
#endif
