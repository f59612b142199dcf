// xml_amend.hpp -- "path/to/node@attr=value" edits of a parsed XML input, used by the
// command-line tools in this directory (the reference's own -a amendments,
// src/input_proc.cpp:218-272, are only reachable through its driver).
#pragma once

#include <sstream>
#include <stdexcept>
#include <string>

#include "pugixml.hpp"

namespace mocc_b200 {
inline void amend_xml(pugi::xml_document &doc, const std::string &spec)
{
    const auto at = spec.find('@');
    const auto eq = spec.find('=', at == std::string::npos ? 0 : at);
    if (at == std::string::npos || eq == std::string::npos)
        throw std::runtime_error("bad amendment (want path/to/node@attr=value): " + spec);
    const std::string path = spec.substr(0, at), attr = spec.substr(at + 1, eq - at - 1), val = spec.substr(eq + 1);
    pugi::xml_node node = doc;
    std::stringstream ss(path);
    std::string part;
    while (std::getline(ss, part, '/')) {
        pugi::xml_node c = node.child(part.c_str());
        node             = c.empty() ? node.append_child(part.c_str()) : c;
    }
    pugi::xml_attribute a = node.attribute(attr.c_str());
    if (a.empty())
        a = node.append_attribute(attr.c_str());
    a.set_value(val.c_str());
}
}
