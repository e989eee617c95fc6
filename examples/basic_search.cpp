// The reference's examples/basic_search.rs, through the C++ host mirror and the B200 path:
// build an index in RAM, search it with a term query, a boolean query and a Count collector.
//   g++ -O2 -std=c++17 examples/basic_search.cpp -Ltantivy_b200/_lib -ltantivy_b200 -Wl,-rpath,$PWD/tantivy_b200/_lib
// (needs a CUDA device at run time: the search path has no CPU fallback)
#include <cstdio>

#include "../tantivy_b200/host/tantivy_host.hpp"

using namespace tantivy_b200;

int main() {
  SchemaBuilder schema_builder;
  const Field title = schema_builder.add_text_field("title", TEXT);
  const Field body = schema_builder.add_text_field("body", TEXT);
  Index index = Index::create_in_ram(schema_builder.build());

  IndexWriter index_writer = index.writer();
  index_writer.add_document(Document().add_text(title, "The Old Man and the Sea")
                                .add_text(body, "He was an old man who fished alone in a skiff in the Gulf Stream and he had gone "
                                                "eighty-four days now without taking a fish."));
  index_writer.add_document(Document().add_text(title, "Of Mice and Men")
                                .add_text(body, "A few miles south of Soledad, the Salinas River drops in close to the hillside bank and "
                                                "runs deep and green. The water is warm too, for it has slipped twinkling over the "
                                                "yellow sands in the sunlight before reaching the narrow pool."));
  index_writer.add_document(Document().add_text(title, "Frankenstein")
                                .add_text(body, "You will rejoice to hear that no disaster has accompanied the commencement of an "
                                                "enterprise which you have regarded with such evil forebodings."));
  index_writer.commit();

  try {
    const Searcher searcher = index.reader().searcher();
    const QueryParser parser = QueryParser::for_index(index, {body});
    for (const char* text : {"sea", "old man", "+old +man", "title:mice the"}) {
      const QueryBox query = parser.parse_query(text);
      const auto top_docs = searcher.search(*query, TopDocs::with_limit(10));
      std::printf("%-16s %zu matching docs, top hits:", text, searcher.search(*query, Count{}));
      for (const auto& hit : top_docs) std::printf("  (%.4f, seg %u doc %u)", hit.first, hit.second.segment_ord, hit.second.doc_id);
      std::printf("\n");
    }
  } catch (const TantivyError& e) {
    std::fprintf(stderr, "search failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
