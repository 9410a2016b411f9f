// TEST INFRASTRUCTURE ONLY — shim for InteractiveComputerGraphics/GenericParameters
// (@42d52ad551fafba600ee99e59fb0f9c7b557e2ed, pinned by the reference in CMake/SetUpExternalProjects.cmake:32-39;
// not vendored in /root/reference).  Written from the call sites in the reference (Simulation.cpp:227-379,
// TimeStep.cpp:41-64, TimeStepDiffDFSPH.cpp:154-281, FluidModel.cpp:120-190): a reflection-style registry of
// named parameters bound to member variables or getter/setter functors.  Carries no hot-path arithmetic.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace GenParam {

class ParameterBase {
 public:
  enum DataTypes { BOOL, DOUBLE, ENUM, FLOAT, FUNCTION, INT8, INT16, INT32, LIST, STRING, STRUCT, UINT8, UINT16, UINT32, VEC_FLOAT, VEC_DOUBLE, VEC_INT32, VEC_UINT32 };
  template <typename T> using GetFunc = std::function<T()>;
  template <typename T> using SetFunc = std::function<void(T)>;
  template <typename T> using GetVecFunc = std::function<T *()>;
  template <typename T> using SetVecFunc = std::function<void(T *)>;

  ParameterBase(const std::string &name, const std::string &label, DataTypes type) : m_name(name), m_label(label), m_type(type) {}
  virtual ~ParameterBase() {}
  const std::string &getName() const { return m_name; }
  const std::string &getLabel() const { return m_label; }
  const std::string &getGroup() const { return m_group; }
  const std::string &getDescription() const { return m_description; }
  const std::string &getHotKey() const { return m_hotKey; }
  DataTypes getType() const { return m_type; }
  bool getReadOnly() const { return m_readOnly; }
  bool getVisible() const { return m_visible; }
  void setName(const std::string &s) { m_name = s; }
  void setLabel(const std::string &s) { m_label = s; }
  void setGroup(const std::string &s) { m_group = s; }
  void setDescription(const std::string &s) { m_description = s; }
  void setHotKey(const std::string &s) { m_hotKey = s; }
  void setReadOnly(bool b) { m_readOnly = b; }
  void setVisible(bool b) { m_visible = b; }
  virtual bool checkType(ParameterBase *other) const { return other->getType() == getType(); }

 protected:
  std::string m_name, m_label, m_group, m_description, m_hotKey;
  DataTypes m_type;
  bool m_readOnly = false, m_visible = true;
};

template <typename T>
class Parameter : public ParameterBase {
 public:
  Parameter(const std::string &name, const std::string &label, DataTypes type, T *ptr)
      : ParameterBase(name, label, type), m_get([ptr]() { return *ptr; }), m_set([ptr](T v) { *ptr = v; }) {}
  Parameter(const std::string &name, const std::string &label, DataTypes type, GetFunc<T> g, SetFunc<T> s)
      : ParameterBase(name, label, type), m_get(g), m_set(s) {}
  virtual ~Parameter() {}
  virtual T getValue() const { return m_get(); }
  virtual void setValue(const T v) {
    if (m_set) m_set(v);
  }
  void setValuePtr(T *ptr) {
    m_get = [ptr]() { return *ptr; };
    m_set = [ptr](T v) { *ptr = v; };
  }

 protected:
  GetFunc<T> m_get;
  SetFunc<T> m_set;
};

template <typename T> struct TypeTag;
template <> struct TypeTag<float> { static const ParameterBase::DataTypes value = ParameterBase::FLOAT, vec = ParameterBase::VEC_FLOAT; };
template <> struct TypeTag<double> { static const ParameterBase::DataTypes value = ParameterBase::DOUBLE, vec = ParameterBase::VEC_DOUBLE; };
template <> struct TypeTag<char> { static const ParameterBase::DataTypes value = ParameterBase::INT8, vec = ParameterBase::VEC_INT32; };
template <> struct TypeTag<short> { static const ParameterBase::DataTypes value = ParameterBase::INT16, vec = ParameterBase::VEC_INT32; };
template <> struct TypeTag<int> { static const ParameterBase::DataTypes value = ParameterBase::INT32, vec = ParameterBase::VEC_INT32; };
template <> struct TypeTag<unsigned char> { static const ParameterBase::DataTypes value = ParameterBase::UINT8, vec = ParameterBase::VEC_UINT32; };
template <> struct TypeTag<unsigned short> { static const ParameterBase::DataTypes value = ParameterBase::UINT16, vec = ParameterBase::VEC_UINT32; };
template <> struct TypeTag<unsigned int> { static const ParameterBase::DataTypes value = ParameterBase::UINT32, vec = ParameterBase::VEC_UINT32; };

template <typename T>
class NumericParameter : public Parameter<T> {
 public:
  NumericParameter(const std::string &name, const std::string &label, T *ptr) : Parameter<T>(name, label, TypeTag<T>::value, ptr) {}
  NumericParameter(const std::string &name, const std::string &label, ParameterBase::GetFunc<T> g, ParameterBase::SetFunc<T> s)
      : Parameter<T>(name, label, TypeTag<T>::value, g, s) {}
  void setMinValue(const T v) { m_min = v; m_hasMin = true; }
  void setMaxValue(const T v) { m_max = v; m_hasMax = true; }
  T getMinValue() const { return m_min; }
  T getMaxValue() const { return m_max; }
  virtual void setValue(const T v) {
    T x = v;
    if (m_hasMin && x < m_min) x = m_min;
    if (m_hasMax && x > m_max) x = m_max;
    Parameter<T>::setValue(x);
  }

 protected:
  T m_min = T(), m_max = T();
  bool m_hasMin = false, m_hasMax = false;
};
using FloatParameter = NumericParameter<float>;
using DoubleParameter = NumericParameter<double>;
using IntParameter = NumericParameter<int>;
using UnsignedIntParameter = NumericParameter<unsigned int>;

class BoolParameter : public Parameter<bool> {
 public:
  BoolParameter(const std::string &name, const std::string &label, bool *ptr) : Parameter<bool>(name, label, ParameterBase::BOOL, ptr) {}
  BoolParameter(const std::string &name, const std::string &label, GetFunc<bool> g, SetFunc<bool> s) : Parameter<bool>(name, label, ParameterBase::BOOL, g, s) {}
};

class StringParameter : public Parameter<std::string> {
 public:
  StringParameter(const std::string &name, const std::string &label, std::string *ptr) : Parameter<std::string>(name, label, ParameterBase::STRING, ptr) {}
  StringParameter(const std::string &name, const std::string &label, GetFunc<std::string> g, SetFunc<std::string> s)
      : Parameter<std::string>(name, label, ParameterBase::STRING, g, s) {}
};

class EnumParameter : public Parameter<int> {
 public:
  struct EnumValue {
    int id;
    std::string name, description;
  };
  EnumParameter(const std::string &name, const std::string &label, int *ptr) : Parameter<int>(name, label, ParameterBase::ENUM, ptr) {}
  EnumParameter(const std::string &name, const std::string &label, GetFunc<int> g, SetFunc<int> s) : Parameter<int>(name, label, ParameterBase::ENUM, g, s) {}
  void addEnumValue(const std::string &name, int &id) {
    id = (int)m_values.size();
    m_values.push_back({id, name, ""});
  }
  void addEnumValue(const std::string &name, const std::string &description, int &id) {
    id = (int)m_values.size();
    m_values.push_back({id, name, description});
  }
  const std::vector<EnumValue> &getEnumValues() const { return m_values; }

 protected:
  std::vector<EnumValue> m_values;
};

template <typename T>
class VectorParameter : public ParameterBase {
 public:
  VectorParameter(const std::string &name, const std::string &label, unsigned int dim, T *ptr)
      : ParameterBase(name, label, TypeTag<T>::vec), m_dim(dim), m_get([ptr]() { return ptr; }), m_set([ptr, dim](T *v) {
          for (unsigned int i = 0; i < dim; i++) ptr[i] = v[i];
        }) {}
  VectorParameter(const std::string &name, const std::string &label, unsigned int dim, GetVecFunc<T> g, SetVecFunc<T> s)
      : ParameterBase(name, label, TypeTag<T>::vec), m_dim(dim), m_get(g), m_set(s) {}
  T *getValue() const { return m_get(); }
  void setValue(T *v) {
    if (m_set) m_set(v);
  }
  unsigned int getDim() const { return m_dim; }

 protected:
  unsigned int m_dim;
  GetVecFunc<T> m_get;
  SetVecFunc<T> m_set;
};
using FloatVectorParameter = VectorParameter<float>;
using DoubleVectorParameter = VectorParameter<double>;

class ParameterObject {
 public:
  using ParameterPtr = std::unique_ptr<ParameterBase>;
  ParameterObject() {}
  virtual ~ParameterObject() {}
  virtual void initParameters() {}
  unsigned int numParameters() const { return (unsigned int)m_parameters.size(); }
  ParameterBase *getParameter(const unsigned int index) { return m_parameters[index].get(); }
  ParameterBase *const getParameter(const unsigned int index) const { return m_parameters[index].get(); }

  template <typename T> int createNumericParameter(const std::string &name, const std::string &label, T *ptr) {
    m_parameters.push_back(ParameterPtr(new NumericParameter<T>(name, label, ptr)));
    return (int)m_parameters.size() - 1;
  }
  template <typename T> int createNumericParameter(const std::string &name, const std::string &label, ParameterBase::GetFunc<T> g, ParameterBase::SetFunc<T> s = {}) {
    m_parameters.push_back(ParameterPtr(new NumericParameter<T>(name, label, g, s)));
    return (int)m_parameters.size() - 1;
  }
  int createBoolParameter(const std::string &name, const std::string &label, bool *ptr) {
    m_parameters.push_back(ParameterPtr(new BoolParameter(name, label, ptr)));
    return (int)m_parameters.size() - 1;
  }
  int createBoolParameter(const std::string &name, const std::string &label, ParameterBase::GetFunc<bool> g, ParameterBase::SetFunc<bool> s = {}) {
    m_parameters.push_back(ParameterPtr(new BoolParameter(name, label, g, s)));
    return (int)m_parameters.size() - 1;
  }
  int createEnumParameter(const std::string &name, const std::string &label, int *ptr) {
    m_parameters.push_back(ParameterPtr(new EnumParameter(name, label, ptr)));
    return (int)m_parameters.size() - 1;
  }
  int createEnumParameter(const std::string &name, const std::string &label, ParameterBase::GetFunc<int> g, ParameterBase::SetFunc<int> s = {}) {
    m_parameters.push_back(ParameterPtr(new EnumParameter(name, label, g, s)));
    return (int)m_parameters.size() - 1;
  }
  int createStringParameter(const std::string &name, const std::string &label, std::string *ptr) {
    m_parameters.push_back(ParameterPtr(new StringParameter(name, label, ptr)));
    return (int)m_parameters.size() - 1;
  }
  int createStringParameter(const std::string &name, const std::string &label, ParameterBase::GetFunc<std::string> g, ParameterBase::SetFunc<std::string> s = {}) {
    m_parameters.push_back(ParameterPtr(new StringParameter(name, label, g, s)));
    return (int)m_parameters.size() - 1;
  }
  template <typename T> int createVectorParameter(const std::string &name, const std::string &label, unsigned int dim, T *ptr) {
    m_parameters.push_back(ParameterPtr(new VectorParameter<T>(name, label, dim, ptr)));
    return (int)m_parameters.size() - 1;
  }
  template <typename T>
  int createVectorParameter(const std::string &name, const std::string &label, unsigned int dim, ParameterBase::GetVecFunc<T> g, ParameterBase::SetVecFunc<T> s = {}) {
    m_parameters.push_back(ParameterPtr(new VectorParameter<T>(name, label, dim, g, s)));
    return (int)m_parameters.size() - 1;
  }

  // convenience setters by parameter id (used as setGroup(ID, "...") throughout the reference)
  void setGroup(const unsigned int id, const std::string &s) { getParameter(id)->setGroup(s); }
  void setDescription(const unsigned int id, const std::string &s) { getParameter(id)->setDescription(s); }
  void setHotKey(const unsigned int id, const std::string &s) { getParameter(id)->setHotKey(s); }
  void setReadOnly(const unsigned int id, bool b) { getParameter(id)->setReadOnly(b); }
  void setVisible(const unsigned int id, bool b) { getParameter(id)->setVisible(b); }
  void setName(const unsigned int id, const std::string &s) { getParameter(id)->setName(s); }
  void setLabel(const unsigned int id, const std::string &s) { getParameter(id)->setLabel(s); }
  std::string getName(const unsigned int id) const { return getParameter(id)->getName(); }
  std::string getLabel(const unsigned int id) const { return getParameter(id)->getLabel(); }
  std::string getGroup(const unsigned int id) const { return getParameter(id)->getGroup(); }
  std::string getDescription(const unsigned int id) const { return getParameter(id)->getDescription(); }
  ParameterBase::DataTypes getType(const unsigned int id) const { return getParameter(id)->getType(); }

  template <typename T> T getValue(const unsigned int id) const { return static_cast<Parameter<T> *>(getParameter(id))->getValue(); }
  template <typename T> void setValue(const unsigned int id, const T v) { static_cast<Parameter<T> *>(getParameter(id))->setValue(v); }
  template <typename T> T *getVecValue(const unsigned int id) const { return static_cast<VectorParameter<T> *>(getParameter(id))->getValue(); }
  template <typename T> void setVecValue(const unsigned int id, T *v) { static_cast<VectorParameter<T> *>(getParameter(id))->setValue(v); }

 protected:
  std::vector<ParameterPtr> m_parameters;
};

}  // namespace GenParam
